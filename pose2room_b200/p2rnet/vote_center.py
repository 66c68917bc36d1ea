"""CenterVoteModule: each seed votes for an object centre (VoteNet-style).

Module surface / state-dict of /root/reference/models/p2rnet/modules/vote_center.py:10-59:
conv_input = 3 x SingleConv (256->256 'cbr', 256->256 'cbr', 256->(3+256)*vote_factor 'c').
Runs on channel-last rows (B*S, 256) with the B200 GEMM / BatchNorm kernels.
"""
import torch
import torch.nn as nn

from .registers import MODULES
from .sub_modules import SingleConv, run_rows


@MODULES.register_module
class CenterVoteModule(nn.Module):
    def __init__(self, cfg, optim_spec=None):
        super().__init__()
        self.optim_spec = optim_spec
        self.origin_joint_id = cfg.dataset_config.origin_joint_id
        self.vote_factor = cfg.config["data"]["vote_factor"]
        self.precision = cfg.config.get("precision", "fp32")
        in_dim = 256
        self.out_dim = in_dim
        self.conv_input = nn.Sequential(
            SingleConv(in_dim, 256, order="cbr"),
            SingleConv(256, 256, order="cbr"),
            SingleConv(256, (3 + self.out_dim) * self.vote_factor, order="c"))

    def forward(self, seed_xyz, seed_features):
        """seed_xyz (B,S,J,3) seed skeletons, seed_features (B,S,256) ->
        vote_xyz (B,S*vf,3), vote_features (B,S*vf,256)."""
        seed_xyz = seed_xyz[:, :, self.origin_joint_id]
        b, s, _ = seed_xyz.shape
        num_vote = s * self.vote_factor
        act = torch.bfloat16 if self.precision == "bf16" else torch.float32
        rows = seed_features.reshape(b * s, -1).to(act)
        net = run_rows(self.conv_input, rows).float()
        net = net.reshape(b, s, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[..., 0:3]).contiguous().reshape(b, num_vote, 3)
        vote_features = (seed_features.unsqueeze(2) + net[..., 3:]).contiguous().reshape(b, num_vote, self.out_dim)
        return vote_xyz, vote_features.contiguous()
