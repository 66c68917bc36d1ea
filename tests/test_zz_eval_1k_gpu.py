"""GPU: BASELINE.json config #5 -- the eval path (decode + far-box filter + 3-D NMS kernels, OBB-IoU kernel, AP) over
1000 synthetic scenes against what the UNMODIFIED reference produced for the same inputs (tests/golden/eval1k.npz;
parse_predictions with scipy Delaunay + numpy NMS, APCalculator with Qhull IoU).

Both tests gate: mAP@0.25 / mAP@0.5 / AR within 5e-3 of the reference's (north_star asks for +-0.1), corner checksums to
fp32-exp rounding, AND selection (pred_mask) bit-exact on all 128 000 proposals with every per-class AP to 1e-6 (green on
a B200 since round 1's end run; the xfail cushion is gone)."""
import time

import numpy as np
import pytest
import torch

from tests import eval1k_helpers as H
from tests.test_geometry_gpu import CFG

pytestmark = pytest.mark.gpu
CHUNK = 125


@pytest.fixture(scope="module")
def run(cuda):
    from pose2room_b200 import ap_helper, synthetic
    g, want_mask = H.load()
    n = int(g["n_scenes"])
    masks, counts, checks, preds, gts = [], [], [], [], []
    t_gpu = 0.0
    for start in range(0, n, CHUNK):
        est, gt = synthetic.make_eval_batch(int(g["seed"]), start, min(CHUNK, n - start))
        est = {k: v.to(cuda) for k, v in est.items()}
        data = {"input_joints": gt["input_joints"].to(cuda)}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eval_dict, parsed = ap_helper.parse_predictions(est, data, CFG)
        eval_dict = ap_helper.assembly_pred_map_cls(eval_dict, parsed, CFG)
        gt_map = ap_helper.assembly_gt_map_cls(ap_helper.parse_groundtruths(gt, CFG))
        t_gpu += time.perf_counter() - t0
        masks.append(eval_dict["pred_mask"])
        counts += [len(x) for x in eval_dict["batch_pred_map_cls"]]
        checks.append(np.abs(parsed["pred_corners_3d"]).sum(axis=(1, 2, 3)))
        preds += eval_dict["batch_pred_map_cls"]
        gts += gt_map
    metrics = {}
    t0 = time.perf_counter()
    for thr in (0.25, 0.5):
        calc = ap_helper.APCalculator(thr)
        calc.step(preds, gts)
        metrics[thr] = calc.compute_metrics()
    t_ap = time.perf_counter() - t0
    print("eval1k: parse+assemble %.2f s (reference: %.0f s), AP x2 %.2f s (reference: %.0f s)" %
          (t_gpu, float(g["parse_seconds"]), t_ap, float(g["ap_seconds_25"]) + float(g["ap_seconds_50"])))
    return g, want_mask, np.concatenate(masks), np.array(counts), np.concatenate(checks), metrics


def _ap_dict(m):
    return {c: m["%d Average Precision" % c] for c in range(22) if "%d Average Precision" % c in m}


def test_map_parity_on_1000_scenes(run):
    g, want_mask, mask, counts, checks, metrics = run
    assert mask.shape == want_mask.shape
    assert (mask != want_mask).mean() < 1e-3, int((mask != want_mask).sum())
    assert abs(int(counts.sum()) - int(g["n_pred"].sum())) <= 1e-3 * int(g["n_pred"].sum())
    assert np.allclose(checks, g["corner_abs_sum"], rtol=1e-5)
    for thr, tag in ((0.25, "25"), (0.5, "50")):
        assert abs(metrics[thr]["mAP"] - float(g["map_" + tag])) < 5e-3, (thr, metrics[thr]["mAP"], float(g["map_" + tag]))
        assert abs(metrics[thr]["AR"] - float(g["ar_" + tag])) < 5e-3, (thr, metrics[thr]["AR"], float(g["ar_" + tag]))


def test_selection_bit_exact_and_per_class_ap_on_1000_scenes(run):
    g, want_mask, mask, counts, checks, metrics = run
    assert np.array_equal(mask, want_mask), int((mask != want_mask).sum())
    assert np.array_equal(counts, g["n_pred"])
    for thr, tag in ((0.25, "25"), (0.5, "50")):
        H.check_ap(_ap_dict(metrics[thr]), metrics[thr]["mAP"], g["ap_" + tag], float(g["map_" + tag]), 1e-6)
