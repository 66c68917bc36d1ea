"""Generate tests/golden/eval1k.npz by EXECUTING the unmodified reference eval path (from /root/reference) on CPU over the
1000-scene evaluation set of BASELINE.json config #5 (pose2room_b200.synthetic.make_eval_scene, seed 2024): the
reference's parse_predictions (scipy Delaunay far-box test, numpy NMS), assembly_*_map_cls and APCalculator (Qhull IoU,
its own Pool(10)) at IoU 0.25 and 0.5.  Build container only:

    python tests/golden/make_golden_eval1k.py [n_scenes]

The inputs are NOT stored -- both sides rebuild them from (seed, scene index); the fixture holds the reference's outputs:
pred_mask (bit-packed), per-scene prediction counts and corner checksums, per-class AP / mAP / AR at both thresholds, and
the same AP numbers for the first SUBSET scenes (what the CPU oracle test can afford)."""
import os.path as osp
import sys
import time

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = osp.dirname(osp.abspath(__file__))

from oracle import ref_import  # noqa: E402
from pose2room_b200 import synthetic  # noqa: E402

SEED, SUBSET, CHUNK = 2024, 40, 50


def metrics(ns, preds, gts, thr):
    calc = ns.ap_helper.APCalculator(thr)
    calc.step(preds, gts)
    m = calc.compute_metrics()
    ap = np.full(22, np.nan)
    rec = np.full(22, np.nan)
    for c in range(22):
        if "%d Average Precision" % c in m:
            ap[c] = m["%d Average Precision" % c]
            rec[c] = m["%d Recall" % c]
    return ap, rec, float(m["mAP"]), float(m["AR"])


def main(n):
    _, cfg = ref_import.build_reference_model(mode="test", joint_num=25, num_frames=256)
    ns = ref_import.import_reference()
    masks, counts, checks, preds, gts = [], [], [], [], []
    t0 = time.time()
    for start in range(0, n, CHUNK):
        est, gt = synthetic.make_eval_batch(SEED, start, min(CHUNK, n - start))
        eval_dict, parsed = ns.ap_helper.parse_predictions(est, {"input_joints": gt["input_joints"]}, cfg.eval_config)
        eval_dict = ns.ap_helper.assembly_pred_map_cls(eval_dict, parsed, cfg.eval_config)
        gt_map = ns.ap_helper.assembly_gt_map_cls(ns.ap_helper.parse_groundtruths(gt, cfg.eval_config))
        masks.append(eval_dict["pred_mask"].astype(np.uint8))
        counts += [len(x) for x in eval_dict["batch_pred_map_cls"]]
        checks.append(np.abs(parsed["pred_corners_3d"]).sum(axis=(1, 2, 3)))
        preds += eval_dict["batch_pred_map_cls"]
        gts += gt_map
        print("parsed %d / %d scenes, %.0f s" % (start + CHUNK, n, time.time() - t0), flush=True)
    out = dict(seed=SEED, n_scenes=n, subset=SUBSET, pred_mask_bits=np.packbits(np.concatenate(masks), axis=1),
               n_pred=np.array(counts, np.int32), corner_abs_sum=np.concatenate(checks),
               parse_seconds=time.time() - t0)
    for thr in (0.25, 0.5):
        tag = "%d" % int(thr * 100)
        t1 = time.time()
        out["ap_" + tag], out["rec_" + tag], out["map_" + tag], out["ar_" + tag] = metrics(ns, preds, gts, thr)
        out["ap_seconds_" + tag] = time.time() - t1
        out["ap_sub_" + tag], _, out["map_sub_" + tag], _ = metrics(ns, preds[:SUBSET], gts[:SUBSET], thr)
        print("IoU %.2f: mAP %.4f AR %.4f (first %d scenes: mAP %.4f), %.0f s" %
              (thr, out["map_" + tag], out["ar_" + tag], SUBSET, out["map_sub_" + tag], time.time() - t1), flush=True)
    np.savez_compressed(osp.join(OUT, "eval1k.npz"), **out)
    print("eval1k.npz written:", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
