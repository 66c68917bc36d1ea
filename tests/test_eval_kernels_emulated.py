"""CPU: the product's eval-path KERNELS (csrc/geometry_ops.cu: decode_boxes, nms3d, box3d_iou -- unmodified, launchers
included) executed by the host emulator (tests/csrc/cuda_emu.h) under the product's own host logic (ap_helper), on the
first `subset` scenes of the 1000-scene evaluation set, against what the UNMODIFIED reference produced for them
(tests/golden/eval1k.npz): far-box / NMS selection bit-exact, corner checksums, per-class AP and mAP at IoU 0.25 / 0.5.
Plus nn_distance and the trajectory-length seed sampling under emulation against the reference goldens (the recorded demo
sequence with its exact ties included).  The same kernels run on the B200 in tests/test_zz_eval_1k_gpu.py /
test_geometry_gpu.py; this is the no-GPU rehearsal of those tests' kernels (host libm instead of the device's)."""
import ctypes
import os.path as osp
import subprocess

import numpy as np
import pytest
import torch

from oracle import geometry_ref as G
from pose2room_b200 import _lib, ap_helper, geometry, synthetic
from tests import eval1k_helpers as H

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
NAMES = ("p2r_decode_boxes", "p2r_nms3d", "p2r_box3d_iou", "p2r_nn_distance", "p2r_uniform_seed_inds")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_eval") / "kernels_emu.so")
    inc = [osp.join(ROOT, "include"), osp.join(ROOT, "pose2room_b200", "csrc"), osp.join(ROOT, "tests", "csrc")]
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-std=c++17", "-w",
                    "-DP2R_HOST_EMULATION"] + sum((["-I", i] for i in inc), []) +
                   [osp.join(ROOT, "tests", "csrc", "kernels_emu.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    for name in NAMES:
        fn = getattr(lib, name)
        fn.argtypes = _lib.SIGNATURES[name]
        fn.restype = ctypes.c_int
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def _p(a):
    return a.ctypes.data


def _emulated_geometry(lib):
    """numpy twins of geometry.decode_boxes / nms3d_batched / box3d_iou_matrix (same marshalling, host arrays)."""
    def decode_boxes(center, log_size, heading_sincos, hip, contact=1.0):
        c = np.ascontiguousarray(center.detach().numpy(), np.float32)
        s = np.ascontiguousarray(log_size.detach().numpy(), np.float32)
        h = np.ascontiguousarray(heading_sincos.detach().numpy(), np.float64)
        hp = np.ascontiguousarray(hip.detach().numpy(), np.float32)
        b, k, t = c.shape[0], c.shape[1], hp.shape[1]
        corners, aabb = np.full((b, k, 8, 3), np.nan), np.full((b, k, 6), np.nan)
        nonempty = np.full((b, k), 7, np.uint8)
        rc = lib.p2r_decode_boxes(_p(c), _p(s), _p(h), _p(hp), 3, b, k, t, float(contact), _p(corners), _p(aabb), _p(nonempty), None)
        assert rc == 0, lib.emu_last_error()
        return torch.from_numpy(corners), torch.from_numpy(aabb), torch.from_numpy(nonempty)

    def nms3d_batched(aabb, score, valid=None, cls=None, thr=0.10, old_type=False):
        a = np.ascontiguousarray(aabb.numpy(), np.float64)
        sc = np.ascontiguousarray(score.numpy(), np.float64)
        v = np.ascontiguousarray(valid.numpy(), np.uint8) if valid is not None else None
        c = np.ascontiguousarray(cls.numpy(), np.int32) if cls is not None else None
        b, k = sc.shape
        keep, order = np.full((b, k), 9, np.uint8), np.full((b, k), -7, np.int32)
        rc = lib.p2r_nms3d(_p(a), _p(sc), _p(v) if v is not None else None, _p(c) if c is not None else None, b, k,
                           float(thr), int(bool(old_type)), _p(keep), _p(order), None)
        assert rc == 0, lib.emu_last_error()
        return torch.from_numpy(keep), torch.from_numpy(order)

    def box3d_iou_matrix(c1, c2):
        a, b = np.ascontiguousarray(c1, np.float64), np.ascontiguousarray(c2, np.float64)
        i3, i2 = np.full((len(a), len(b)), np.nan), np.full((len(a), len(b)), np.nan)
        rc = lib.p2r_box3d_iou(_p(a), _p(b), len(a), len(b), _p(i3), _p(i2), None)
        assert rc == 0, lib.emu_last_error()
        return torch.from_numpy(i3), torch.from_numpy(i2)
    return decode_boxes, nms3d_batched, box3d_iou_matrix


def test_eval_kernels_under_emulation_reproduce_the_reference_on_the_eval_subset(emu, monkeypatch):
    from tests.test_geometry_gpu import CFG
    g, want_mask = H.load()
    n = int(g["subset"])
    dec, nms, iou = _emulated_geometry(emu)
    monkeypatch.setattr(geometry, "decode_boxes", dec)
    monkeypatch.setattr(geometry, "nms3d_batched", nms)
    monkeypatch.setattr(geometry, "box3d_iou_matrix", iou)
    est, gt = synthetic.make_eval_batch(int(g["seed"]), 0, n)
    eval_dict, parsed = ap_helper.parse_predictions(est, {"input_joints": gt["input_joints"]}, CFG)   # the product's host logic
    assert np.array_equal(eval_dict["pred_mask"], want_mask[:n])
    assert np.allclose(np.abs(parsed["pred_corners_3d"]).sum(axis=(1, 2, 3)), g["corner_abs_sum"][:n], rtol=1e-6)
    eval_dict = ap_helper.assembly_pred_map_cls(eval_dict, parsed, CFG)
    assert [len(x) for x in eval_dict["batch_pred_map_cls"]] == g["n_pred"][:n].tolist()
    gt_map = []
    for i in range(n):
        cur = []
        for j in range(10):
            if gt["box_label_mask"][i, j] == 1:
                hs = gt["heading"][i, j].numpy()
                cur.append((int(gt["sem_cls_label"][i, j]),
                            G.get_3d_box(np.exp(gt["size"][i, j].numpy()), np.arctan2(hs[0], hs[1]), gt["center_label"][i, j].numpy())))
        gt_map.append(cur)
    for thr in (0.25, 0.5):
        tag = "%d" % int(thr * 100)
        calc = ap_helper.APCalculator(thr)
        calc.step(eval_dict["batch_pred_map_cls"], gt_map)
        m = calc.compute_metrics()
        ap = {c: m["%d Average Precision" % c] for c in range(22) if "%d Average Precision" % c in m}
        H.check_ap(ap, m["mAP"], g["ap_sub_" + tag], float(g["map_sub_" + tag]), 1e-9)


def test_nn_distance_kernel_under_emulation_is_bit_exact_vs_reference_goldens(emu, golden_geometry):
    g = golden_geometry
    for case in ("demo", "a", "b"):
        pc1, pc2 = np.ascontiguousarray(g["nnd_%s_pc1" % case]), np.ascontiguousarray(g["nnd_%s_pc2" % case])
        b, n, c = pc1.shape
        m = pc2.shape[1]
        for mode, tag in ((0, "l2"), (2, "l1s"), (1, "l1")):
            d1, i1 = np.full((b, n), np.nan, np.float32), np.full((b, n), -1, np.int64)
            d2, i2 = np.full((b, m), np.nan, np.float32), np.full((b, m), -1, np.int64)
            assert emu.p2r_nn_distance(_p(pc1), _p(pc2), b, n, m, c, mode, 1.0, _p(d1), _p(i1), _p(d2), _p(i2), None) == 0
            for key, got in (("d1", d1), ("i1", i1), ("d2", d2), ("i2", i2)):
                assert np.array_equal(got, g["nnd_%s_%s_%s" % (case, tag, key)]), (case, tag, key)


def test_seed_sampling_kernel_under_emulation_on_the_recorded_demo_sequence(emu):
    """427 repeated frames -> exact ties in |cum - target|: the kernel must pick the reference's (first) frame every time."""
    g = np.load(osp.join(ROOT, "tests", "golden", "demo.npz"))
    joints = np.ascontiguousarray(g["raw_joints"][g["frame_ids"]][None], np.float32)          # (1, 768, 53, 3)
    seeds = np.full((1, 512), -1, np.int64)
    assert emu.p2r_uniform_seed_inds(_p(joints), 53 * 3, 1, 768, 512, _p(seeds), None) == 0   # hip = joint 0, read in place
    assert np.array_equal(seeds, g["gen_seed_inds"])


def test_nms_and_iou_kernels_under_emulation_vs_reference_goldens(emu, golden_geometry):
    """Rehearsal of test_geometry_gpu.py's NMS / IoU checks: the reference's own picks (plain, old-type, same-class, 2-D)
    on its fixture boxes, random sets against the oracle, pairwise oriented-box IoU against the reference's Qhull values."""
    g = golden_geometry
    _, nms, iou = _emulated_geometry(emu)

    def picks(boxes, thr, old_type=False, with_cls=False):
        boxes = np.asarray(boxes, np.float64)
        t = torch.from_numpy(boxes)
        cls = t[None, :, 7].to(torch.int32) if with_cls else None
        keep, order = nms(t[None, :, 0:6].contiguous(), t[None, :, 6].contiguous(), None, cls, thr, old_type)
        return [int(i) for i in order[0].numpy() if i >= 0]

    for t in range(4):
        boxes = g["nms%d_boxes" % t]
        assert picks(boxes[:, :7], 0.10) == g["nms%d_pick" % t].tolist()
        assert picks(boxes[:, :7], 0.25, old_type=True) == g["nms%d_pick_old" % t].tolist()
        assert picks(boxes, 0.10, with_cls=True) == g["nms%d_pick_cls" % t].tolist()
        k = boxes.shape[0]
        b3 = np.zeros((k, 7))
        b3[:, 0:2], b3[:, 3:5], b3[:, 5], b3[:, 6] = boxes[:, 0:2], boxes[:, 3:5], 1.0, boxes[:, 6]   # geometry.nms_2d_faster
        assert picks(b3, 0.10) == g["nms%d_pick_2d" % t].tolist()
    rng = np.random.default_rng(5)
    for k in (1, 2, 31, 128, 200):
        lo = rng.normal(0, 1.0, size=(k, 3))
        boxes = np.concatenate([lo, lo + rng.uniform(0.05, 1.5, size=(k, 3)), rng.uniform(size=(k, 1))], 1)
        assert picks(boxes, 0.1) == G.nms_3d_faster(boxes, 0.1)
    i3, i2 = iou(g["box_corners"], g["box_corners"])
    ok = ~np.isnan(g["box_iou3d"]) & ~np.eye(24, dtype=bool)
    assert np.allclose(i3.numpy()[ok], g["box_iou3d"][ok], atol=1e-9) and np.allclose(i2.numpy()[ok], g["box_iou2d"][ok], atol=1e-9)


def test_knn_and_graph_offset_kernels_under_emulation_vs_reference_goldens(emu, golden_pointnet2):
    """Rehearsal of test_geometry_gpu.py::test_knn_vs_oracle_and_reference_golden."""
    from oracle.pointnet2_ref import knn_ref
    for name in ("p2r_knn_graph", "p2r_graph_offset"):
        fn = getattr(emu, name)
        fn.argtypes = _lib.SIGNATURES[name]
        fn.restype = ctypes.c_int
    g = golden_pointnet2
    x = np.ascontiguousarray(g["knn_x"], np.float32)
    b, c, n = x.shape
    idx = np.full((b, n, 8), -1, np.int64)
    assert emu.p2r_knn_graph(_p(x), b, c, n, 8, _p(idx), None) == 0
    assert np.array_equal(np.sort(idx, -1), np.sort(g["knn_idx"], -1))                         # the reference's torch.topk sets
    assert np.array_equal(np.sort(idx, -1), np.sort(knn_ref(torch.from_numpy(x), 8).numpy(), -1))
    ref_idx = np.ascontiguousarray(g["knn_idx"], np.int64)
    off = np.full(g["knn_offset"].shape, np.nan, np.float32)
    assert emu.p2r_graph_offset(_p(x), _p(ref_idx), b, c, n, 8, _p(off), None) == 0
    assert np.array_equal(off, g["knn_offset"])
    # live configuration of the backbone: (B,3,T) hip trajectory, k = 20
    rng = np.random.default_rng(0)
    xt = np.cumsum(rng.normal(0, 0.05, size=(1, 3, 1024)), axis=2).astype(np.float32)
    got = np.full((1, 1024, 20), -1, np.int64)
    assert emu.p2r_knn_graph(_p(xt), 1, 3, 1024, 20, _p(got), None) == 0
    want = knn_ref(torch.from_numpy(xt), 20).numpy()
    assert (np.sort(got, -1) == np.sort(want, -1)).mean() > 0.999 and (got[..., 0] == np.arange(1024)[None]).all()


@pytest.mark.parametrize("V", [6, 25])        # (53, the reference rig, passes too: 70 s under the emulator)
def test_graph_conv_weight_kernels_under_emulation_match_the_einsum_formulation(emu, V):
    """csrc/graph_conv.cu: W_eff = sum_k W_k (x) A_k, b_eff and the fold of their gradients back onto the conv weight,
    conv bias and A (ref: ConvTemporalGraphical.forward, stgcn_layers.py:58-67 -- conv 64 -> K*64 then
    einsum('nkctv,kvw->nctw')), against numpy einsums in float64."""
    for name in ("p2r_gcn_build_weight", "p2r_gcn_reduce_weight_grad"):
        fn = getattr(emu, name)
        fn.argtypes = _lib.SIGNATURES[name]
        fn.restype = ctypes.c_int
    rng = np.random.default_rng(V)
    K, Cc = 11, 64
    W = (rng.normal(size=(K * Cc, Cc)) / 8).astype(np.float32)
    bias = rng.normal(size=K * Cc).astype(np.float32)
    A = (rng.random((K, V, V)) * (rng.random((K, V, V)) < 0.15)).astype(np.float32)          # sparse, like adjacency * importance
    w_eff, w_eff_t = np.zeros((V * Cc, V * Cc), np.uint16), np.zeros((V * Cc, V * Cc), np.uint16)
    b_eff = np.full(V * Cc, np.nan, np.float32)
    assert emu.p2r_gcn_build_weight(_p(W), _p(bias), _p(A), K, V, Cc, Cc, _p(w_eff), _p(w_eff_t), _p(b_eff), None) == 0
    Wk = W.reshape(K, Cc, Cc).astype(np.float64)
    want = np.einsum("kvw,koi->wovi", A.astype(np.float64), Wk).reshape(V * Cc, V * Cc)
    got = (w_eff.astype(np.uint32) << 16).view(np.float32)
    assert np.abs(got - want).max() <= 4e-3 * np.abs(want).max()                            # stored as bf16
    assert np.array_equal(w_eff_t, w_eff.T)
    want_b = np.einsum("ko,kw->wo", bias.reshape(K, Cc).astype(np.float64), A.astype(np.float64).sum(1)).reshape(-1)
    assert np.abs(b_eff - want_b).max() <= 1e-5 * np.abs(want_b).max()
    # backward fold
    dWe = rng.normal(size=(V * Cc, V * Cc)).astype(np.float32)
    dbe = rng.normal(size=V * Cc).astype(np.float32)
    dW, db, dA = np.full((K * Cc, Cc), np.nan, np.float32), np.full(K * Cc, np.nan, np.float32), np.full((K, V, V), np.nan, np.float32)
    assert emu.p2r_gcn_reduce_weight_grad(_p(dWe), _p(dbe), _p(W), _p(bias), _p(A), K, V, Cc, Cc, _p(dW), _p(db), _p(dA), None) == 0
    d4 = dWe.reshape(V, Cc, V, Cc).astype(np.float64)                                        # [w, co, v, ci]
    want_dW = np.einsum("kvw,wovi->koi", A.astype(np.float64), d4).reshape(K * Cc, Cc)
    want_db = np.einsum("kw,wo->ko", A.astype(np.float64).sum(1), dbe.reshape(V, Cc).astype(np.float64)).reshape(-1)
    want_dA = np.einsum("koi,wovi->kvw", Wk, d4) + np.einsum("ko,wo->kw", bias.reshape(K, Cc).astype(np.float64),
                                                           dbe.reshape(V, Cc).astype(np.float64))[:, None, :]
    want_dA = np.where(A != 0, want_dA, 0.0)                                                 # dA is 0 where A == 0
    for name, a, b in (("dW", dW, want_dW), ("db", db, want_db), ("dA", dA, want_dA)):
        assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max(), (name, np.abs(a - b).max(), np.abs(b).max())
