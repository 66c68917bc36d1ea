"""BASELINE.json config #0 -- the reference's own demo sequence through the reference's own demo path, on CPU:
demo/inputs/input_joints_1.npy (341 recorded frames, the 53-joint rig) -> Demo_DataSet.__getitem__'s frame resampling
(demo.py:33-48: np.linspace(...).round().astype(uint16), float32 cast) -> P2RNet.generate(data, eval=False) as
demo.predict calls it (demo.py:260-266), with the test YAML's shape (768 frames, 512 seeds, 128 proposals).  The pretrained
weight is an external download (README.md:37), so weights are the deterministic synthetic ones of the other goldens.

    python tests/golden/make_golden_demo.py        (build container only)

The recorded input travels inside the fixture (it is the one real-data fixture the reference ships for this path, 217 KB);
tests/golden/demo.npz also holds the resampled frame ids and the reference's outputs."""
import os.path as osp
import sys

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = osp.dirname(osp.abspath(__file__))

from oracle import pointnet2_ref, ref_import  # noqa: E402
from pose2room_b200 import synthetic  # noqa: E402
from tests.golden.make_golden_model import _cfg_for, _rebuild  # noqa: E402

T, J, S, P = 768, 53, 512, 128
EP_KEYS = ["seed_inds", "aggregated_vote_inds", "vote_xyz", "aggregated_vote_xyz", "center", "size", "heading",
           "objectness_scores", "sem_cls_scores"]


def main():
    raw = np.load(osp.join(ref_import.REF_ROOT, "demo", "inputs", "input_joints_1.npy"))
    frame_ids = np.linspace(0, raw.shape[0] - 1, T).round().astype(np.uint16)           # demo.py:43
    data = {"input_joints": torch.from_numpy(raw[frame_ids].astype(np.float32))[None], "sample_idx": ["input_joints_1"]}
    net, cfg = _rebuild(_cfg_for("test", J, T, S, P), J, "test")
    sd = synthetic.deterministic_state_dict(net.state_dict(), seed=7)
    net.load_state_dict(sd)
    net.eval()
    with torch.no_grad():
        ep, eval_dict, parsed = net.generate(data, eval=False)
    out = {"raw_joints": raw, "frame_ids": frame_ids}
    for k in sd:
        if k.endswith(".mdn.mu"):
            out["mu_" + k] = sd[k].numpy()
    for k in EP_KEYS:
        out["gen_" + k] = ep[k].numpy()
    out["gen_pred_mask"] = eval_dict["pred_mask"]
    out["gen_corners"] = parsed["pred_corners_3d"]
    out["gen_obj_prob"] = parsed["obj_prob"]
    out["gen_pred_sem_cls"] = parsed["pred_sem_cls"]
    out["gen_npred"] = np.array([len(x) for x in eval_dict["batch_pred_map_cls"]])
    np.savez_compressed(osp.join(OUT, "demo.npz"), **out)
    print("demo.npz: kept", int(eval_dict["pred_mask"].sum()), "of", P, "proposals;",
          "distinct seed frames", len(np.unique(ep["seed_inds"].numpy())), "; duplicate input frames",
          T - len(np.unique(frame_ids)))


if __name__ == "__main__":
    main()
