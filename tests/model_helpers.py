"""Shared helpers for the model-level tests: golden configs, deterministic weights, product / oracle builders."""
import os.path as osp

import numpy as np
import torch

from pose2room_b200 import synthetic
from pose2room_b200.config import P2RConfig

GOLDEN = osp.join(osp.dirname(osp.abspath(__file__)), "golden", "p2rnet.npz")
CONFIGS = {"small": (2, 160, 25, 64, 16), "ref53": (1, 96, 53, 128, 32), "bl": (1, 1024, 25, 512, 128)}
EP_KEYS = ["seed_inds", "aggregated_vote_inds", "vote_xyz", "aggregated_vote_xyz", "center", "size", "heading",
           "objectness_scores", "sem_cls_scores"]


def load_golden():
    return np.load(GOLDEN)


def make_cfg(name, mode, precision="fp32"):
    B, T, J, S, P = CONFIGS[name]
    return P2RConfig(mode=mode, joint_num=J, num_frames=T, precision=precision, num_seeds=S, num_target=P)


def make_product(name, mode, golden, precision="fp32", train_noise_off=True):
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(0)
    np.random.seed(0)
    net = P2RNet(make_cfg(name, mode, precision))
    net.load_state_dict(weights_for(name, net.state_dict(), golden, train_noise_off))
    return net


def weights_for(name, template_sd, golden, train_noise_off=True):
    sd = synthetic.deterministic_state_dict(template_sd, seed=7)
    for k in sd:
        if k.endswith(".mdn.mu"):
            sd[k] = torch.from_numpy(golden["%s_%s" % (name, k)])
        if k.endswith(".mdn.log_sigma") and train_noise_off:
            sd[k] = torch.full_like(sd[k], -50.0)
    return sd


def make_data(name, device=None):
    B, T, J, S, P = CONFIGS[name]
    data = synthetic.make_batch(B, T, J, seed=1234)
    if device is not None:
        data = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
    return data
