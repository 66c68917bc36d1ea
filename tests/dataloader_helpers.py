"""Shared helpers for the sample -> batch (dataloader) tests: read tests/golden/dataloader.npz."""
import os.path as osp

import numpy as np

GOLDEN = osp.join(osp.dirname(osp.abspath(__file__)), "golden", "dataloader.npz")
KEYS = ["input_joints", "box_label_mask", "sem_cls_label", "center_label", "size", "heading",
        "vote_label", "vote_label_mask"]
ROT_ANGLES = [-np.pi, -0.5 * np.pi, 0, 0.5 * np.pi]


def load():
    return np.load(GOLDEN)


def raw_sample(g, name):
    n = len(g["%s_class_id" % name])
    nodes = [dict(class_id=g["%s_class_id" % name][i], centroid=g["%s_centroid" % name][i],
                  R_mat=g["%s_R_mat" % name][i], size=g["%s_size" % name][i]) for i in range(n)]
    return dict(skeleton_joints=g["%s_joints" % name], skeleton_joint_votes=g["%s_votes" % name],
                object_nodes=nodes, name=str(name))


def cases(g):
    """Yield (sample name, case tag, draws or None, num_frames)."""
    for name in g["names"]:
        name = str(name)
        nf = int(g["%s_num_frames" % name])
        yield name, "noaug", None, nf
        for i, d in enumerate(g["%s_draws" % name]):
            yield name, "aug%d" % i, (int(d[0]), ROT_ANGLES[int(d[1])], float(d[2])), nf


def assert_same(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype, "%s: dtype %s != %s" % (what, got.dtype, want.dtype)
    assert got.shape == want.shape, "%s: shape %s != %s" % (what, got.shape, want.shape)
    assert np.array_equal(got, want), "%s: %d of %d entries differ (max |d| = %g)" % (
        what, int((got != want).sum()), got.size, float(np.abs(got.astype(np.float64) - want).max()))


class Cfg:
    """Minimal stand-in for the reference's CONFIG as far as the dataset class reads it."""

    def __init__(self, num_frames, batch_size=2, no_height=True, distributed=False):
        self.config = {"data": {"dataset": "virtualhome", "split": "/nonexistent", "num_frames": num_frames,
                                "no_height": no_height, "max_gt_boxes": 10},
                       "device": {"distributed": distributed, "gpu": 0, "num_workers": 0},
                       "train": {"batch_size": batch_size}, "val": {"batch_size": batch_size},
                       "test": {"batch_size": 1}}
        self.dataset_config = None


def dataset_for(store, num_frames, use_height=False, aug=False, device=None):
    from pose2room_b200 import dataloader as DL
    return DL.P2RNet_VirtualHome(Cfg(num_frames, no_height=not use_height), "train" if aug else "test",
                                 packed=store, device=device)


def host_side(ds, indices, draws):
    """The host half of make_batch (parameter blocks + box labels) without the launch."""
    return ds.host_side(indices, draws)
