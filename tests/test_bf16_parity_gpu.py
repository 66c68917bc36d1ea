"""GPU: evidence that the BENCHED mode (bf16 activations / tcgen05 GEMMs, fp32 accumulation, fp32 BatchNorm statistics,
fp32 parameters, losses and geometry) computes the same model as the fp32 parity mode -- the mode that is held to 1e-4 of
the unmodified reference's goldens in test_model_gpu.py.  Same weights, same inputs, same process:

  1. layer-wise error budget: relative L2 error of every ST-GCN block output, the seed features and the votes, bf16 vs
     fp32, at the BASELINE shape (T = 1024, J = 25): <= 1e-2 each (bf16 has 8 mantissa bits: 2^-9 = 2e-3 per rounding);
  2. detection-level agreement through the eval path (generate -> decode -> far-box -> NMS -> per-class AP): the bf16
     model's detections scored against the fp32 model's detections as ground truth, mAP@0.25 / 0.5 within 0.1 (north_star's
     accuracy tolerance) of the fp32 model scored against itself, and >= 90 % of the fp32 boxes recovered at IoU 0.25;
  3. a 200-step training A/B from the same initial weights on the same batches: the two loss curves stay together.
"""
import numpy as np
import pytest
import torch

from pose2room_b200 import synthetic
from pose2room_b200.config import P2RConfig

pytestmark = pytest.mark.gpu


def _net(precision, mode, dev, T=1024, J=25, **kw):
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(0)
    np.random.seed(0)
    net = P2RNet(P2RConfig(mode=mode, joint_num=J, num_frames=T, precision=precision, **kw))
    net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7))
    return net.to(dev)


def _to(data, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in data.items()}


def _rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm())


def test_layerwise_error_budget_bf16_vs_fp32(cuda):
    from pose2room_b200 import gemm_sm100
    from pose2room_b200.p2rnet import stgcn
    data = _to(synthetic.make_batch(2, 1024, 25, seed=1234), cuda)
    taps = {}
    original = stgcn.st_gcn_block.forward_rows

    def tapped(self, x, A, sparsity=None):
        out = original(self, x, A, sparsity)
        taps[self._tag].append(out.detach().float().cpu())
        return out
    stgcn.st_gcn_block.forward_rows = tapped
    try:
        results = {}
        for precision in ("fp32", "bf16"):
            if precision == "bf16":
                gemm_sm100.install()
            try:
                net = _net(precision, "train", cuda).train()
                for i, blk in enumerate(net.backbone.st_gcn_networks):
                    blk._tag = "block%d" % i
                    taps[blk._tag] = []
                with torch.no_grad():
                    torch.manual_seed(5)
                    ep = net(data)
                results[precision] = dict({k: v[0] for k, v in taps.items()},
                                          **{k: ep[k].detach().float().cpu() for k in ("seed_features", "vote_xyz", "vote_features")},
                                          seed_inds=ep["seed_inds"].cpu())
            finally:
                if precision == "bf16":
                    gemm_sm100.uninstall()
    finally:
        stgcn.st_gcn_block.forward_rows = original
    assert torch.equal(results["fp32"]["seed_inds"], results["bf16"]["seed_inds"])
    errs = {k: _rel_l2(results["bf16"][k], results["fp32"][k]) for k in results["fp32"] if k != "seed_inds"}
    print("bf16 vs fp32 relative L2 per layer:", {k: round(v, 5) for k, v in errs.items()})
    for k, e in errs.items():
        assert e <= 1e-2, (k, e, errs)


def _detections(net, batches, cfg):
    """-> per scene: list of (class, corners (8,3), score) for every kept box (its arg-max class and objectness), plus the
    per-class-proposal prediction list the reference's AP consumes (ap_helper.py:294-350)."""
    boxes, preds = [], []
    with torch.no_grad():
        for data in batches:
            ep, eval_dict, parsed = net.generate(data, eval=False)
            preds += eval_dict["batch_pred_map_cls"]
            for b in range(eval_dict["pred_mask"].shape[0]):
                keep = np.nonzero(eval_dict["pred_mask"][b])[0]
                keep = [k for k in keep if parsed["obj_prob"][b, k] > cfg["conf_thresh"]]
                boxes.append([(int(parsed["pred_sem_cls"][b, k]), parsed["pred_corners_3d"][b, k], float(parsed["obj_prob"][b, k]))
                              for k in keep])
    return boxes, preds


def test_bf16_detections_agree_with_fp32_detections(cuda):
    from pose2room_b200 import ap_helper, gemm_sm100, geometry
    batches = [_to(synthetic.make_batch(16, 1024, 25, seed=4000 + i), cuda) for i in range(4)]     # 64 scenes
    out = {}
    for precision in ("fp32", "bf16"):
        if precision == "bf16":
            gemm_sm100.install()
        try:
            net = _net(precision, "test", cuda).eval()
            out[precision] = _detections(net, batches, net.cfg.eval_config)
        finally:
            if precision == "bf16":
                gemm_sm100.uninstall()
    gt = [[(c, corners) for c, corners, _ in scene] for scene in out["fp32"][0]]
    n_gt = sum(len(s) for s in gt)
    assert n_gt >= 64, n_gt                      # the fixture's weights leave several boxes per scene after NMS
    scores = {}
    for precision in ("fp32", "bf16"):
        for thr in (0.25, 0.5):
            calc = ap_helper.APCalculator(thr)
            calc.step(out[precision][1], gt)
            scores[precision, thr] = calc.compute_metrics()["mAP"]
    # box-level recall: share of the fp32 model's kept boxes that the bf16 model also keeps (IoU >= 0.25 / 0.5, any class)
    hit = {0.25: 0, 0.5: 0}
    for a, b in zip(out["fp32"][0], out["bf16"][0]):
        if not a:
            continue
        if not b:
            continue
        iou, _ = geometry.box3d_iou_matrix(np.stack([x[1] for x in a]), np.stack([x[1] for x in b]))
        best = iou.max(dim=1).values.cpu().numpy()
        for thr in hit:
            hit[thr] += int((best >= thr).sum())
    recall = {thr: hit[thr] / float(n_gt) for thr in hit}
    print("fp32 boxes %d, bf16 boxes %d; mAP (fp32 detections as ground truth): %s; recall of fp32 boxes: %s" % (
        n_gt, sum(len(s) for s in out["bf16"][0]), {"%s@%.2f" % k: round(v, 4) for k, v in scores.items()}, recall))
    for thr in (0.25, 0.5):
        assert abs(scores["bf16", thr] - scores["fp32", thr]) <= 0.1, (thr, scores)
    assert recall[0.25] >= 0.9, recall


def test_loss_curve_ab_200_steps(cuda):
    """Same initial weights, same 8 batches in the same order, same RNG seed for the sampled mixture heads, AdamW 1e-3,
    200 steps in each mode (eager, like the reference's trainer).  FPS picks and ReLU masks differ between the modes after
    the first update, so the curves are compared as curves: mean over windows of 20 steps."""
    from pose2room_b200 import gemm_sm100
    T, J, S, P, B, steps = 512, 25, 256, 64, 8, 200
    pool = [_to(synthetic.make_batch(B, T, J, seed=7000 + i), cuda) for i in range(8)]
    curves = {}
    for precision in ("fp32", "bf16"):
        if precision == "bf16":
            gemm_sm100.install()
        try:
            net = _net(precision, "train", cuda, T=T, J=J, num_seeds=S, num_target=P).train()
            opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3)
            torch.manual_seed(99)
            losses = []
            for i in range(steps):
                opt.zero_grad(set_to_none=True)
                data = pool[i % len(pool)]
                loss = net.loss(net(data), data)["total"]
                loss.backward()
                opt.step()
                losses.append(loss.detach())
            curves[precision] = torch.stack(losses).double().cpu().numpy()
        finally:
            if precision == "bf16":
                gemm_sm100.uninstall()
    a, b = curves["fp32"], curves["bf16"]
    assert np.isfinite(a).all() and np.isfinite(b).all()
    wa, wb = a.reshape(-1, 20).mean(1), b.reshape(-1, 20).mean(1)
    print("loss curve, mean of 20-step windows  fp32:", np.round(wa, 3).tolist(), " bf16:", np.round(wb, 3).tolist())
    assert abs(a[0] - b[0]) <= 0.01 * abs(a[0]), (a[0], b[0])            # first step: same weights, forward error only
    assert wa[-1] < 0.8 * wa[0] and wb[-1] < 0.8 * wb[0], (wa, wb)        # both actually train
    assert np.abs(wb - wa).max() <= 0.08 * np.abs(wa).max(), (wa, wb)     # and stay together window by window
