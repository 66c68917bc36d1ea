"""CPU: the arithmetic of the fused detection-loss kernel (csrc/loss_math.h, compiled for the host by this test through
tests/csrc/detection_loss_host.c, which mirrors the kernel's indexing) against the CPU oracle's loss
(oracle/model_ref.py RefP2RNet.loss -- pinned by the unmodified reference's goldens in test_model_oracle.py): the ten
reported numbers, every label / assignment decision (through the numbers that depend on them) and the gradient of
`total` with respect to all six differentiable inputs.  The CUDA kernel itself is covered by test_model_gpu.py."""
import ctypes
import os.path as osp
import subprocess
import types

import numpy as np
import pytest
import torch

from oracle.model_ref import RefP2RNet
from pose2room_b200 import synthetic

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
KEYS32 = ["vote_loss", "objectness_loss", "center_loss", "size_loss", "sem_cls_loss", "pos_ratio", "neg_ratio", "obj_acc"]


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("loss") / "detection_loss_host.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", osp.join(ROOT, "pose2room_b200", "csrc"),
                    osp.join(ROOT, "tests", "csrc", "detection_loss_host.c"), "-o", so, "-lm"], check=True)
    return ctypes.CDLL(so)


def make_case(seed, B=3, T=48, J=25, S=20, P=16, C=22, heading_dtype=torch.float64, prefix_masks=True):
    """Predictions scattered around the ground truth so that near / far / in-between proposals, masked and unmasked
    seeds, |error| on both sides of the huber knee and every class all occur."""
    g = torch.Generator().manual_seed(seed)
    gt = synthetic.make_batch(B, T, J, seed=seed)
    if not prefix_masks:     # the reference's compacted-index quirk (loss.py:128 vs :68) only shows with holes
        gt["box_label_mask"][:, 0] = 0.0
        gt["box_label_mask"][:, 3] = 1.0
    joints = gt["input_joints"]
    seed_inds = torch.sort(torch.stack([torch.randperm(T, generator=g)[:S] for _ in range(B)]), dim=1)[0]
    skeleton = torch.gather(joints, 1, seed_inds.view(B, S, 1, 1).expand(B, S, J, 3)).contiguous()
    which = torch.randint(0, 10, (B, P), generator=g)
    near = torch.gather(gt["center_label"], 1, which.unsqueeze(-1).expand(B, P, 3))
    spread = torch.tensor([0.05, 0.25, 0.45, 1.5])[torch.randint(0, 4, (B, P), generator=g)].unsqueeze(-1)
    agg = near + spread * torch.randn(B, P, 3, generator=g)
    sem_obj = 3.0 * torch.randn(B, P, 2 + C, generator=g)
    est = {
        "seed_skeleton": skeleton, "seed_inds": seed_inds,
        "vote_xyz": skeleton[:, :, 0] + 0.8 * torch.randn(B, S, 3, generator=g),
        "aggregated_vote_xyz": agg,
        "center": agg + 0.3 * torch.randn(B, P, 3, generator=g),
        "size": torch.gather(gt["size"], 1, which.unsqueeze(-1).expand(B, P, 3)) + 0.9 * torch.randn(B, P, 3, generator=g),
        "heading": (1.2 * torch.randn(B, P, 2, generator=g)).to(heading_dtype),
        "objectness_scores": sem_obj[..., 0:2], "sem_cls_scores": sem_obj[..., 2:],
    }
    return est, gt, sem_obj


def oracle_loss(est, gt, sem_obj):
    leaves = {k: est[k].clone().requires_grad_(True) for k in ("vote_xyz", "center", "size", "heading")}
    so = sem_obj.clone().requires_grad_(True)
    e = dict(est, **leaves)
    e["objectness_scores"], e["sem_cls_scores"] = so[..., 0:2], so[..., 2:]
    me = types.SimpleNamespace(o=0, _nn_distance=RefP2RNet._nn_distance, _huber=RefP2RNet._huber)
    out = RefP2RNet.loss(me, e, gt)
    out["total"].backward()
    grads = {k: v.grad for k, v in leaves.items()}
    grads["objectness_scores"], grads["sem_cls_scores"] = so.grad[..., 0:2], so.grad[..., 2:]
    return out, grads


def host_loss(lib, est, gt, sem_obj, g_total=1.0):
    B, S, J = est["seed_skeleton"].shape[:3]
    T, P, G, C = gt["vote_label"].shape[1], est["center"].shape[1], gt["center_label"].shape[1], sem_obj.shape[2] - 2
    h64 = est["heading"].dtype == torch.float64
    arr = lambda t, dt: np.ascontiguousarray(t.detach().numpy().astype(dt, copy=False))
    so = arr(sem_obj, np.float32)                                   # obj / sem are row-strided slices of this
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ins = [arr(est["vote_xyz"], np.float32), arr(est["center"], np.float32), arr(est["size"], np.float32),
           arr(est["heading"], np.float64 if h64 else np.float32)]
    tail = [arr(est["aggregated_vote_xyz"], np.float32), arr(est["seed_skeleton"], np.float32),
            arr(est["seed_inds"], np.int64), arr(gt["vote_label"], np.float32), arr(gt["vote_label_mask"], np.int64),
            arr(gt["center_label"], np.float32), arr(gt["box_label_mask"], np.float32), arr(gt["size"], np.float32),
            arr(gt["heading"], np.float32), arr(gt["sem_cls_label"], np.int64)]
    out32, out64, scales = np.zeros(8, np.float32), np.zeros(2, np.float64), np.zeros(4, np.float64)
    u = [np.zeros((B, S, 3), np.float32), np.zeros((B, P, 3), np.float32), np.zeros((B, P, 3), np.float32),
         np.zeros((B, P, 3), np.float32), np.zeros((B, P, 2), np.float64 if h64 else np.float32),
         np.zeros((B, P, 2), np.float32), np.zeros((B, P, C), np.float32)]
    obj_ptr = ctypes.c_void_p(so.ctypes.data)
    sem_ptr = ctypes.c_void_p(so.ctypes.data + 2 * 4)
    lib.host_detection_loss(*[ptr(a) for a in ins], int(h64), obj_ptr, 2 + C, sem_ptr, 2 + C, *[ptr(a) for a in tail],
                            B, S, J, T, P, G, C, 0, ptr(out32), ptr(out64), ptr(scales), *[ptr(a) for a in u])
    g32, g64 = np.zeros(8, np.float32), np.array([0.0, g_total])
    d = [np.zeros_like(a) for a in (u[0], u[1], u[3], u[4], u[5], u[6])]
    lib.host_detection_loss_grad(ptr(g32), ptr(g64), ptr(scales), *[ptr(a) for a in u[:5]], int(h64), ptr(u[5]), ptr(u[6]),
                                 B, S, P, C, *[ptr(a) for a in d])
    names = ["vote_xyz", "center", "size", "heading", "objectness_scores", "sem_cls_scores"]
    return out32, out64, dict(zip(names, d))


@pytest.mark.parametrize("seed,heading_dtype,prefix", [(1, torch.float64, True), (2, torch.float64, True),
                                                       (3, torch.float32, True), (4, torch.float64, False)])
def test_ten_numbers_and_gradients_match_the_oracle(host_lib, seed, heading_dtype, prefix):
    est, gt, sem_obj = make_case(seed, heading_dtype=heading_dtype, prefix_masks=prefix)
    want, want_grads = oracle_loss(est, gt, sem_obj)
    out32, out64, grads = host_loss(host_lib, est, gt, sem_obj)
    # the case exercises every branch
    assert 0.05 < float(want["pos_ratio"]) < 0.95 and float(want["neg_ratio"]) > 0.02
    assert 0.0 < float(want["obj_acc"]) < 1.0 and float(want["vote_loss"].detach()) > 0.0
    for k, v in zip(KEYS32, out32):
        assert abs(float(v) - float(want[k])) <= 2e-6 * max(1.0, abs(float(want[k]))), (k, float(v), float(want[k]))
    assert abs(out64[0] - float(want["heading_loss"])) <= 2e-6 * max(1.0, abs(float(want["heading_loss"])))
    assert abs(out64[1] - float(want["total"])) <= 2e-6 * abs(float(want["total"]))
    for k, gref in want_grads.items():
        gref = gref.numpy().astype(np.float64)
        err = np.abs(grads[k].astype(np.float64) - gref).max()
        assert err <= 2e-6 * max(1e-3, np.abs(gref).max()), (k, err, np.abs(gref).max())
        assert np.abs(gref).max() > 0


def test_upstream_gradients_of_single_terms(host_lib):
    """What `(3 * loss['size_loss'] + 0.5 * loss['heading_loss']).backward()` must produce: `total` weighs both terms by
    10 (loss.py:168) and the backward is linear in the upstream gradients (p2rl_term_weights)."""
    est, gt, sem_obj = make_case(5)
    leaves = {k: est[k].clone().requires_grad_(True) for k in ("vote_xyz", "center", "size", "heading")}
    me = types.SimpleNamespace(o=0, _nn_distance=RefP2RNet._nn_distance, _huber=RefP2RNet._huber)
    out = RefP2RNet.loss(me, dict(est, **leaves), gt)
    (3.0 * out["size_loss"] + 0.5 * out["heading_loss"]).backward()
    _, _, from_total = host_loss(host_lib, est, gt, sem_obj, g_total=1.0)
    want_size = leaves["size"].grad.numpy() / 3.0 * 10.0
    want_head = leaves["heading"].grad.numpy() / 0.5 * 10.0
    assert np.abs(from_total["size"] - want_size).max() <= 2e-6 * np.abs(want_size).max()
    assert np.abs(from_total["heading"] - want_head).max() <= 2e-6 * np.abs(want_head).max()
    _, _, none = host_loss(host_lib, est, gt, sem_obj, g_total=0.0)
    assert all(np.abs(v).max() == 0.0 for v in none.values())


def test_sample_without_valid_box_is_far_from_everything(host_lib):
    """The reference raises on a sample with no valid box (min over an empty dim, loss.py:128); the kernel instead labels
    its proposals negative and keeps going -- a padded sample must not poison the batch."""
    est, gt, sem_obj = make_case(6)
    gt["box_label_mask"][1] = 0.0
    out32, out64, grads = host_loss(host_lib, est, gt, sem_obj)
    assert np.isfinite(out32).all() and np.isfinite(out64).all()
    assert np.abs(grads["size"][1]).max() == 0.0 and np.abs(grads["sem_cls_scores"][1]).max() == 0.0
    assert np.abs(grads["objectness_scores"][1]).max() > 0.0          # all negatives, all unmasked
