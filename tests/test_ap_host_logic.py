"""CPU: host logic of the AP evaluation core (pose2room_b200/ap_helper.py: same-scene IoU blocks computed chunk by
chunk, greedy matching, VOC AP) with the GPU IoU launch replaced by the numpy oracle -- the chunking must not change a
single TP / FP decision.  The IoU kernel itself is covered by tests/test_geometry_gpu.py."""
import numpy as np
import torch

from oracle import geometry_ref as G
from pose2room_b200 import ap_helper


def _scenes(rng, n_scenes, n_cls=4):
    preds, gts = [], []
    for _ in range(n_scenes):
        n_gt = int(rng.integers(0, 5))
        gl, pl = [], []
        for _ in range(n_gt):
            c, s, th = rng.normal(0, 1.0, 3), rng.uniform(0.3, 1.5, 3), rng.uniform(-np.pi, np.pi)
            cls = int(rng.integers(0, n_cls))
            gl.append((cls, G.get_3d_box(s, th, c)))
            for _ in range(int(rng.integers(0, 3))):          # detections near the box, some good, some poor
                noise = rng.choice([0.05, 0.4])
                pl.append((cls if rng.random() < 0.8 else int(rng.integers(0, n_cls)),
                           G.get_3d_box(s * rng.uniform(0.8, 1.2, 3), th + rng.normal(0, noise), c + rng.normal(0, noise, 3)),
                           float(rng.random())))
        for _ in range(int(rng.integers(0, 3))):              # stray detections (also in scenes without GT)
            pl.append((int(rng.integers(0, n_cls)), G.get_3d_box(rng.uniform(0.3, 1.5, 3), rng.uniform(-3, 3), rng.normal(0, 2.0, 3)),
                       float(rng.random())))
        preds.append(pl)
        gts.append(gl)
    return preds, gts


def _oracle_iou_matrix(a, b):
    m = np.array([[G.box3d_iou(x, y)[0] for y in b] for x in a], dtype=np.float64).reshape(len(a), len(b))
    return torch.from_numpy(m), None


def test_chunked_same_scene_iou_gives_the_oracle_ap(monkeypatch):
    rng = np.random.default_rng(4)
    preds, gts = _scenes(rng, 60)
    want = G.eval_map(preds, gts, 0.25)
    calls = []

    def fake(a, b):
        calls.append((len(a), len(b)))
        return _oracle_iou_matrix(a, b)
    monkeypatch.setattr(ap_helper.geometry, "box3d_iou_matrix", fake)
    results = []
    for limit in (1 << 22, 40, 1):                              # one chunk / a few scenes per chunk / one scene per chunk
        monkeypatch.setattr(ap_helper, "_IOU_CHUNK_ENTRIES", limit)
        calls.clear()
        calc = ap_helper.APCalculator(0.25)
        calc.step(preds[:25], gts[:25])
        calc.step(preds[25:], gts[25:])
        results.append((calc.compute_metrics(), list(calls)))
    for metrics, _ in results:
        assert metrics == results[0][0]                         # chunking changes nothing, bit for bit
    assert len(results[0][1]) < len(results[1][1]) < len(results[2][1])
    got = results[0][0]
    want_ap, want_map = want
    assert abs(got["mAP"] - want_map) < 1e-12 and 0.05 < want_map < 0.95      # a non-trivial evaluation
    for cls, ap in want_ap.items():
        assert abs(got["%s Average Precision" % cls] - ap) < 1e-12, cls


def test_scene_blocks_cover_exactly_the_same_scene_pairs(monkeypatch):
    monkeypatch.setattr(ap_helper.geometry, "box3d_iou_matrix", _oracle_iou_matrix)
    monkeypatch.setattr(ap_helper, "_IOU_CHUNK_ENTRIES", 12)
    rng = np.random.default_rng(9)
    box = lambda: G.get_3d_box(rng.uniform(0.3, 1.5, 3), rng.uniform(-3, 3), rng.normal(0, 0.5, 3))
    pred = {0: [(box(), 0.9), (box(), 0.1)], 1: [], 2: [(box(), 0.5)], 5: [(box(), 0.4), (box(), 0.3), (box(), 0.2)]}
    gt = {0: [box(), box(), box()], 1: [box()], 2: [], 5: [box()], 7: [box()]}
    blocks = ap_helper._same_scene_iou_rows(pred, gt)
    assert sorted(blocks) == [0, 5] and blocks[0].shape == (2, 3) and blocks[5].shape == (3, 1)
    for img, m in blocks.items():
        for i, (b, _) in enumerate(pred[img]):
            for j, g in enumerate(gt[img]):
                assert m[i, j] == G.box3d_iou(b, g)[0]
