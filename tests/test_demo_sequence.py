"""BASELINE.json config #0: the reference's recorded demo sequence (demo/inputs/input_joints_1.npy: 341 frames of the
53-joint rig) through the demo path -- frame resampling of Demo_DataSet (demo.py:33-48) and P2RNet.generate as
demo.predict calls it (demo.py:260-266) at the test YAML's shape (768 frames, 512 seeds, 128 proposals) -- against what
the UNMODIFIED reference produced (tests/golden/demo.npz, made by tests/golden/make_golden_demo.py).

Real data is a tie stress the synthetic sets are not: up-sampling 341 frames to 768 repeats 427 of them, so the
trajectory-length seed sampling sees exact ties (224 distinct seed frames among 512 seeds) and FPS / ball query run on
duplicated points.  CPU: the oracle port.  GPU: the product (non-gating for one round: first run is the round-end one)."""
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


def _compare(ep, pred_mask, corners, g, float_tol):
    for k in EP_KEYS:
        want, got = g["gen_" + k], ep[k].detach().cpu().numpy()
        if want.dtype.kind in "iu":
            assert np.array_equal(got, want), k
        else:
            tol = max(float_tol, 2e-5 * float(np.abs(want).max())) if k.endswith("_scores") else float_tol
            assert float(np.abs(got - want).max()) <= tol, (k, float(np.abs(got - want).max()))
    assert np.array_equal(pred_mask, g["gen_pred_mask"])
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


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="new shape (1 x 768 x 53) and a real-data tie stress for kernels that are green on the "
                                        "synthetic goldens; added without a GPU at hand, non-gating until its first run")
def test_product_on_the_demo_sequence_vs_reference(cuda):
    g = np.load(GOLDEN)
    net = _product("test")
    net.load_state_dict(_weights(g, net.state_dict()))
    net = net.to(cuda).eval()
    data = _inputs(g)
    data["input_joints"] = data["input_joints"].to(cuda)
    with torch.no_grad():
        ep, eval_dict, parsed = net.generate(data, eval=False)
    _compare(ep, eval_dict["pred_mask"], parsed["pred_corners_3d"], g, 1e-4)
    assert [len(x) for x in eval_dict["batch_pred_map_cls"]] == g["gen_npred"].tolist()
