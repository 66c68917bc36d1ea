"""Generate tests/golden/dataloader.npz by EXECUTING the unmodified reference dataset class
(/root/reference/models/p2rnet/dataloader.py: P2RNet_VirtualHome.__getitem__ / augment_data /
collate_fn) on CPU.  Build container only.

    python tests/golden/make_golden_dataloader.py

h5py is absent from this image, so `h5py.File` is replaced by an in-memory stand-in that hands the
reference the same objects an HDF5 file written by utils/tools.py:109-139 would (float32 datasets,
int32 class ids, a group of nodes keyed '0','1',... in h5py's alphabetical key order).  Everything
after the file read -- the RNG draws, the flip / rotate / translate arithmetic with its
float32 / float64 mix, rot2head, log-size, frame picking -- is the reference's own code.

Stored per case: the raw sample (inputs), the RNG seed, the draws the reference made with that
seed, and every array of the returned dict.
"""
import json
import os
import os.path as osp
import random
import sys
import tempfile

import numpy as np

ROOT = osp.dirname(osp.dirname(osp.dirname(osp.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = osp.dirname(osp.abspath(__file__))

from oracle import dataloader_ref, ref_import  # noqa: E402
from pose2room_b200 import synthetic  # noqa: E402

_STORE = {}


class _Dataset:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, key):  # h5py reads return fresh arrays (augment_data edits them in place)
        return np.array(self.arr[key], copy=True)


class _Group(dict):
    def keys(self):  # h5py iterates names alphabetically
        return sorted(dict.keys(self))


class _File(_Group):
    def __init__(self, path, mode="r"):
        s = _STORE[path]
        nodes = _Group()
        for i, n in enumerate(s["object_nodes"]):
            nodes[str(i)] = {"class_id": _Dataset(np.array([n["class_id"]], np.int32)),
                             "centroid": _Dataset(n["centroid"]), "R_mat": _Dataset(n["R_mat"]),
                             "size": _Dataset(n["size"])}
        super().__init__(skeleton_joints=_Dataset(s["skeleton_joints"]),
                         skeleton_joint_votes=_Dataset(s["skeleton_joint_votes"]), object_nodes=nodes)

    def close(self):
        pass


class _Cfg:
    def __init__(self, config):
        self.config = config
        self.dataset_config = None


def main():
    ns = ref_import.import_reference()
    import h5py
    h5py.File = _File
    import models.p2rnet.dataloader as ref_dl

    rng = np.random.default_rng(2024)
    # (name, raw frames, joints, network frames): down-sampling, up-sampling (num_frames > raw), 53-joint rig
    specs = [("s0", 40, 25, 16), ("s1", 23, 25, 32), ("s2", 64, 25, 16), ("s3", 31, 53, 12)]
    out = {"names": np.array([s[0] for s in specs])}
    split_dir = tempfile.mkdtemp(prefix="p2r_split_")
    paths = []
    for name, F, J, _ in specs:
        path = "/mem/%s.hdf5" % name
        s = synthetic.make_raw_sample(rng, F, J, name=name)
        if name == "s1":  # a single-box sample (the minimum the loss supports)
            s["object_nodes"] = s["object_nodes"][:1]
        _STORE[path] = s
        paths.append(path)
        out["%s_joints" % name] = s["skeleton_joints"]
        out["%s_votes" % name] = s["skeleton_joint_votes"]
        out["%s_class_id" % name] = np.array([n["class_id"] for n in s["object_nodes"]], np.int32)
        out["%s_centroid" % name] = np.stack([n["centroid"] for n in s["object_nodes"]])
        out["%s_R_mat" % name] = np.stack([n["R_mat"] for n in s["object_nodes"]])
        out["%s_size" % name] = np.stack([n["size"] for n in s["object_nodes"]])
    for mode in ("train", "test"):
        json.dump(paths, open(osp.join(split_dir, mode + ".json"), "w"))

    keys = ["input_joints", "box_label_mask", "sem_cls_label", "center_label", "size", "heading",
            "vote_label", "vote_label_mask"]
    for si, (name, F, J, nf) in enumerate(specs):
        config = {"data": {"split": split_dir, "num_frames": nf, "no_height": True, "max_gt_boxes": 10}}
        test_ds = ref_dl.P2RNet_VirtualHome(_Cfg(config), "test")
        item = test_ds[si]
        assert item["sample_idx"] == name
        for k in keys:
            out["%s_noaug_%s" % (name, k)] = item[k]
        out["%s_num_frames" % name] = np.int64(nf)
        train_ds = ref_dl.P2RNet_VirtualHome(_Cfg(config), "train")
        # walk seeds until every (flip, angle) combination has been seen once
        seen, seed, seeds, draws = set(), 0, [], []
        while len(seen) < 8:
            random.seed(seed)
            np.random.seed(seed)
            d = dataloader_ref.draw_augmentation(random, np.random)
            combo = (d[0], dataloader_ref.ROT_ANGLES.index(d[1]))
            if combo not in seen:
                seen.add(combo)
                random.seed(seed)
                np.random.seed(seed)
                item = train_ds[si]
                for k in keys:
                    out["%s_aug%d_%s" % (name, len(seeds), k)] = item[k]
                seeds.append(seed)
                draws.append([d[0], combo[1], d[2]])
            seed += 1
        out["%s_seeds" % name] = np.array(seeds, np.int64)
        out["%s_draws" % name] = np.array(draws, np.float64)  # (flip, angle index, offset scale)

    # use_height branch (dead in the shipped YAMLs: no_height True), one train + one test case
    config = {"data": {"split": split_dir, "num_frames": 16, "no_height": False, "max_gt_boxes": 10}}
    out["s0_height_noaug_input_joints"] = ref_dl.P2RNet_VirtualHome(_Cfg(config), "test")[0]["input_joints"]
    random.seed(3)
    np.random.seed(3)
    out["s0_height_aug_input_joints"] = ref_dl.P2RNet_VirtualHome(_Cfg(config), "train")[0]["input_joints"]
    out["s0_height_aug_seed"] = np.int64(3)

    # collate_fn on two items of equal shape
    config = {"data": {"split": split_dir, "num_frames": 16, "no_height": True, "max_gt_boxes": 10}}
    ds = ref_dl.P2RNet_VirtualHome(_Cfg(config), "test")
    batch = ref_dl.collate_fn([ds[0], ds[2]])
    assert batch["sample_idx"] == ["s0", "s2"]
    for k in keys:
        out["collate_%s" % k] = batch[k].numpy()

    path = osp.join(OUT, "dataloader.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6), len(out), "arrays")


if __name__ == "__main__":
    main()
