"""BASELINE.json config #0: the reference's recorded demo sequence (demo/inputs/input_joints_1.npy: 341 frames of the
53-joint rig) through the demo path -- frame resampling of Demo_DataSet (demo.py:33-48) and P2RNet.generate as
demo.predict calls it (demo.py:260-266) at the test YAML's shape (768 frames, 512 seeds, 128 proposals) -- against what
the UNMODIFIED reference produced (tests/golden/demo.npz, made by tests/golden/make_golden_demo.py).

Real data is a tie stress the synthetic sets are not: up-sampling 341 frames to 768 repeats 427 of them, so the
trajectory-length seed sampling sees exact ties (224 distinct seed frames among 512 seeds) and FPS / ball query run on
duplicated points.  CPU: the oracle port.  GPU: the product, gating.

NMS selection and exact score ties.  While the person stands still, clusters of proposals pool the SAME 16 votes and get
bit-identical objectness (the golden holds a 14-way and an 8-way exact tie among its 128 scores, all within 0.457..0.463).
The reference orders the scores with `np.argsort(score)` (net_utils/nms.py:51), which is not a stable sort: which member of
a tie comes first is implementation-defined (numpy 2.3's AVX-512 sort, which made the golden, orders the 14-way tie
33,32,31,30,34,35,64,66,60,62,63,36,65,67; the reference's pinned numpy 1.19 introsort orders it differently; the kernel
takes the higher index first).  So `pred_mask` is compared exactly OUTSIDE tie classes and by the number
of survivors INSIDE each class of tied scores -- the strongest statement the reference itself supports (round 1's B200
failure of this test was exactly that: 4 of 128 bits, two swaps inside those two classes, reproduced on the CPU by feeding
the golden's own logits to the emulated kernel)."""
import os.path as osp

import numpy as np
import pytest
import torch

from pose2room_b200 import synthetic
from pose2room_b200.config import P2RConfig

GOLDEN = osp.join(osp.dirname(osp.abspath(__file__)), "golden", "demo.npz")
T, J, S, P = 768, 53, 512, 128
EP_KEYS = ["seed_inds", "aggregated_vote_inds", "vote_xyz", "aggregated_vote_xyz", "center", "size", "heading",
           "objectness_scores", "sem_cls_scores"]


def _inputs(g):
    raw = g["raw_joints"]
    ids = np.linspace(0, raw.shape[0] - 1, T).round().astype(np.uint16)       # demo.py:43 (the make_batch kernel's frame
    assert np.array_equal(ids, g["frame_ids"])                                # picking is held to this in test_dataloader_math)
    return {"input_joints": torch.from_numpy(raw[ids].astype(np.float32))[None], "sample_idx": ["input_joints_1"]}


def _weights(g, template):
    sd = synthetic.deterministic_state_dict(template, seed=7)
    for k in sd:
        if k.endswith(".mdn.mu"):
            sd[k] = torch.from_numpy(g["mu_" + k])
    return sd


def _product(mode="test"):
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(0)
    np.random.seed(0)
    return P2RNet(P2RConfig(mode=mode, joint_num=J, num_frames=T, num_seeds=S, num_target=P))


def tie_classes(score, tol):
    """Indices grouped by chains of scores closer than `tol` (tol = 0: exactly tied scores)."""
    order = np.argsort(score, kind="stable")
    cls, cur = [], [int(order[0])]
    for a, b in zip(order[:-1], order[1:]):
        if score[b] - score[a] <= tol:
            cur.append(int(b))
        else:
            cls.append(cur)
            cur = [int(b)]
    cls.append(cur)
    return cls


def assert_mask_equal_up_to_ties(got, want, score, tol):
    """got / want (K,) 0/1: equal wherever a score is unique, same number of survivors inside every tie class."""
    for members in tie_classes(np.asarray(score, np.float64), tol):
        if len(members) == 1:
            assert got[members[0]] == want[members[0]], ("unique score, different decision", members[0])
        else:
            assert int(got[members].sum()) == int(want[members].sum()), ("tie class", members, got[members], want[members])


def _compare(ep, pred_mask, corners, g, float_tol, score_tol=None):
    for k in EP_KEYS:
        want, got = g["gen_" + k], ep[k].detach().cpu().numpy()
        if want.dtype.kind in "iu":
            assert np.array_equal(got, want), k
        else:
            tol = max(float_tol, 2e-5 * float(np.abs(want).max())) if k.endswith("_scores") else float_tol
            assert float(np.abs(got - want).max()) <= tol, (k, float(np.abs(got - want).max()))
    if score_tol is None:       # same numpy, same scores bit for bit (the oracle here): exact
        assert np.array_equal(pred_mask, g["gen_pred_mask"])
    else:
        for b in range(pred_mask.shape[0]):
            assert_mask_equal_up_to_ties(pred_mask[b], g["gen_pred_mask"][b], g["gen_obj_prob"][b], score_tol)
        assert int(pred_mask.sum()) == int(g["gen_pred_mask"].sum())
    assert float(np.abs(corners - g["gen_corners"]).max()) < 1e-4


def test_fixture_is_the_tie_stress_it_claims():
    g = np.load(GOLDEN)
    assert g["raw_joints"].shape == (341, 53, 3) and len(np.unique(g["frame_ids"])) == 341
    assert len(np.unique(g["gen_seed_inds"])) < S // 2 and 1 <= int(g["gen_pred_mask"].sum()) < P


def test_oracle_on_the_demo_sequence_vs_reference():
    from oracle.model_ref import RefP2RNet
    g = np.load(GOLDEN)
    net = RefP2RNet(_weights(g, _product("train").state_dict()), joint_num=J, num_seeds=S, num_target=P, training=False)
    ep, parsed = net.generate(_inputs(g))
    _compare(ep, parsed["pred_mask"], parsed["corners"], g, 2e-5)


def test_fixture_holds_exact_score_ties_and_the_mask_is_tie_order_dependent():
    """The premise of the tie-aware comparison, checked on the CPU: the golden's scores hold exact ties (one class of at
    least 8), the oracle with this container's numpy reproduces the golden mask bit for bit, and the comparison rejects a
    flipped decision on a proposal whose score is unique."""
    from oracle import geometry_ref as G
    g = np.load(GOLDEN)
    score = g["gen_obj_prob"][0].astype(np.float64)
    classes = [c for c in tie_classes(score, 0.0) if len(c) > 1]
    assert max(len(c) for c in classes) >= 8
    hip = _inputs(g)["input_joints"][:, :, 0].numpy()
    parsed = G.parse_predictions(g["gen_center"], g["gen_size"], g["gen_heading"], g["gen_objectness_scores"],
                                 g["gen_sem_cls_scores"], hip)
    assert np.array_equal(parsed["pred_mask"], g["gen_pred_mask"])
    want = g["gen_pred_mask"][0]
    assert_mask_equal_up_to_ties(want, want, score, 0.0)
    flipped = want.copy()
    unique = [c[0] for c in tie_classes(score, 0.0) if len(c) == 1]
    flipped[unique[0]] ^= 1
    with pytest.raises(AssertionError):
        assert_mask_equal_up_to_ties(flipped, want, score, 0.0)


@pytest.mark.gpu
def test_eval_kernels_on_the_goldens_own_outputs(cuda):
    """Teacher-forced: the reference's own network outputs (exact ties included) through the product's decode / far-box /
    NMS kernels -> corners to 1e-6, selection equal up to the order inside classes of EXACTLY tied scores."""
    from pose2room_b200 import ap_helper
    from tests.test_geometry_gpu import CFG
    g = np.load(GOLDEN)
    est = {k: torch.from_numpy(g["gen_" + k]).to(cuda) for k in ["center", "size", "heading", "objectness_scores", "sem_cls_scores"]}
    data = {"input_joints": _inputs(g)["input_joints"].to(cuda)}
    eval_dict, parsed = ap_helper.parse_predictions(est, data, CFG)
    assert float(np.abs(parsed["pred_corners_3d"] - g["gen_corners"]).max()) < 1e-6
    assert np.array_equal(parsed["obj_prob"], g["gen_obj_prob"]) or float(np.abs(parsed["obj_prob"] - g["gen_obj_prob"]).max()) < 1e-7
    assert_mask_equal_up_to_ties(eval_dict["pred_mask"][0], g["gen_pred_mask"][0], g["gen_obj_prob"][0], 0.0)
    assert int(eval_dict["pred_mask"].sum()) == int(g["gen_pred_mask"].sum())


@pytest.mark.gpu
def test_product_on_the_demo_sequence_vs_reference(cuda):
    g = np.load(GOLDEN)
    net = _product("test")
    net.load_state_dict(_weights(g, net.state_dict()))
    net = net.to(cuda).eval()
    data = _inputs(g)
    data["input_joints"] = data["input_joints"].to(cuda)
    with torch.no_grad():
        ep, eval_dict, parsed = net.generate(data, eval=False)
    # scores within 2e-6 of each other count as tied: the product's logits are within 1e-6 of the reference's (fp32 GEMM
    # summation order), which is 30 ulps of a probability near 0.46 and can reorder neighbours that are not exact ties
    _compare(ep, eval_dict["pred_mask"], parsed["pred_corners_3d"], g, 1e-4, score_tol=2e-6)
    assert [len(x) for x in eval_dict["batch_pred_map_cls"]] == g["gen_npred"].tolist()
