"""Stand-alone configuration objects with the shape the reference's modules expect.

The reference passes a `CONFIG` object around (configs/config_utils.py:46-160) from which the hot path
reads only: cfg.config (the YAML as a nested dict), cfg.dataset_config (joint_num, origin_joint_id,
num_class, contact_dist_thresh; configs/dataset_config.py:15-16,56,78) and, outside training,
cfg.eval_config (the YAML's `test:` block + dataset_config; config_utils.py:140-160).  When running
inside the reference, its own CONFIG works unchanged; this module builds an equivalent object without the
reference tree (GPU box, bench, tests).  Defaults = configs/config_files/p2rnet_{train,test}.yaml.
"""
import copy


class DatasetConfig:
    """Subset of configs/dataset_config.py:Dataset_Config used by the hot path."""

    def __init__(self, joint_num=53, origin_joint_id=0, num_class=22, contact_dist_thresh=1.0):
        self.joint_num = joint_num
        self.origin_joint_id = origin_joint_id
        self.num_class = num_class
        self.contact_dist_thresh = contact_dist_thresh
        self.num_heading_bin = 12


_DEFAULT = {
    "method": "P2RNet",
    "seed": 42,
    "device": {"use_gpu": True, "gpu": 0, "distributed": False, "is_main_process": True},
    "data": {"dataset": "virtualhome", "num_frames": 768, "num_seeds": 512, "seed_sampling": "uniform",
             "max_gt_boxes": 10, "num_target": 128, "vote_factor": 1, "cluster_sampling": "vote_fps",
             "no_height": True, "num_gaussian": 100},
    "model": {"backbone": {"method": "STGCN", "loss": "Null"},
              "centervoting": {"method": "CenterVoteModule", "loss": "Null"},
              "detection": {"method": "ProposalNet", "loss": "BoxNetDetectionLoss"}},
    "optimizer": {"method": "Adam", "lr": 1e-3, "betas": [0.9, 0.999], "eps": 1e-8, "weight_decay": 0,
                  "clip_norm": -1},
    "train": {"epochs": 180, "phase": "full", "freeze": [], "batch_size": 8},
    "val": {"phase": "full", "batch_size": 8},
    "test": {"phase": "full", "batch_size": 1, "use_cls_nms": False, "use_3d_nms": True,
             "ap_iou_thresholds": [0.25, 0.5], "remove_far_box": True, "nms_iou": 0.10,
             "use_old_type_nms": False, "per_class_proposal": True, "conf_thresh": 0.05, "multi_mode": False,
             "sample_cls": False},
    "demo": {"phase": "full"},
}


class P2RConfig:
    """Drop-in for the reference's CONFIG as far as P2RNet / its modules / its loss are concerned."""

    def __init__(self, mode="train", joint_num=53, num_frames=768, precision="fp32", **data_overrides):
        self.config = copy.deepcopy(_DEFAULT)
        self.config["mode"] = mode
        self.config["data"]["num_frames"] = num_frames
        self.config["data"].update(data_overrides)
        self.config["precision"] = precision
        self.dataset_config = DatasetConfig(joint_num=joint_num)
        if mode != "train":
            self.config.setdefault("train", {"freeze": []})
        # mount_external_config (config_utils.py:140-160)
        self.eval_config = dict(self.config["test"])
        self.eval_config["dataset_config"] = self.dataset_config
        self.eval_config["cls_nms"] = self.eval_config["use_cls_nms"]

    def log_string(self, *args, **kwargs):
        pass
