"""CenterVoteModule: each seed votes for an object centre (VoteNet-style).

Module surface / state-dict of /root/reference/models/p2rnet/modules/vote_center.py:10-59:
conv_input = 3 x SingleConv (256->256 'cbr', 256->256 'cbr', 256->(3+256)*vote_factor 'c').
Runs on channel-last rows (B*S, 256) with the B200 GEMM / BatchNorm kernels.
"""
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib
from .registers import MODULES
from .sub_modules import SingleConv, run_rows


# residual adds + L2 normalisation of the vote features as one forward and one backward launch (csrc/vote_ops.cu).
# Default since round 2 (parity-green on a B200 with the flag on); P2R_FUSED_VOTE=0 selects the torch path it is tested
# against.  Read at call time (a test / bench child flips it per process).
def fused_vote_enabled():
    return os.environ.get("P2R_FUSED_VOTE", "1") != "0"


class _VoteTail(Function):
    """net [R, 3+C] (f32 / bf16), seed_xyz (B,S,3) f32 (rows uniformly strided), seed_features (B,S,C) f32 ->
    vote_xyz (B,S,3), L2-normalised vote_features (B,S,C)   (vote_center.py:52-58 + network.py:89-90)."""

    @staticmethod
    def forward(ctx, net, seed_xyz, seed_features):
        b, s, c = seed_features.shape
        rows = b * s
        net = net if net.is_contiguous() else net.contiguous()
        if net.dtype not in (torch.float32, torch.bfloat16):
            net = net.float()
        sf = seed_features.float().contiguous()
        if seed_xyz.dtype != torch.float32 or seed_xyz.stride(2) != 1 or seed_xyz.stride(0) != s * seed_xyz.stride(1):
            seed_xyz = seed_xyz.float().contiguous()
        dev = net.device
        vote_xyz = torch.empty(b, s, 3, dtype=torch.float32, device=dev)
        vote_feat = torch.empty(b, s, c, dtype=torch.float32, device=dev)
        norm = torch.empty(rows, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("p2r_vote_tail", net.data_ptr(), int(net.dtype == torch.bfloat16), seed_xyz.data_ptr(),
                      int(seed_xyz.stride(1)), sf.data_ptr(), rows, c, vote_xyz.data_ptr(), vote_feat.data_ptr(),
                      norm.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(vote_feat, norm)
        ctx.net_meta = (net.dtype, rows, c, (b, s))
        return vote_xyz, vote_feat

    @staticmethod
    def backward(ctx, g_xyz, g_feat):
        vote_feat, norm = ctx.saved_tensors
        dtype, rows, c, (b, s) = ctx.net_meta
        dev = vote_feat.device
        g_xyz = g_xyz.float().contiguous() if g_xyz is not None else None
        g_feat = g_feat.float().contiguous() if g_feat is not None else None
        d_net = torch.empty(rows, 3 + c, dtype=dtype, device=dev)
        d_sf = torch.empty(b, s, c, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("p2r_vote_tail_grad", g_xyz.data_ptr() if g_xyz is not None else None,
                      g_feat.data_ptr() if g_feat is not None else None, vote_feat.data_ptr(), norm.data_ptr(), rows, c,
                      d_net.data_ptr(), int(dtype == torch.bfloat16), d_sf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return d_net, None, d_sf


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

    def forward(self, seed_xyz, seed_features, normalize=False):
        """seed_xyz (B,S,J,3) seed skeletons, seed_features (B,S,256) ->
        vote_xyz (B,S*vf,3), vote_features (B,S*vf,256).
        normalize=True (not in the reference's signature; passed by P2RNet when the fused vote tail is enabled) also
        applies the L2 normalisation of network.py:89-90, so the tail runs as one kernel."""
        seed_xyz = seed_xyz[:, :, self.origin_joint_id]
        b, s, _ = seed_xyz.shape
        num_vote = s * self.vote_factor
        act = torch.bfloat16 if self.precision == "bf16" else torch.float32
        rows = seed_features.reshape(b * s, -1).to(act)
        if normalize and self.vote_factor == 1 and not seed_xyz.requires_grad:
            return _VoteTail.apply(run_rows(self.conv_input, rows), seed_xyz, seed_features)
        net = run_rows(self.conv_input, rows).float()
        net = net.reshape(b, s, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[..., 0:3]).contiguous().reshape(b, num_vote, 3)
        vote_features = (seed_features.unsqueeze(2) + net[..., 3:]).contiguous().reshape(b, num_vote, self.out_dim)
        vote_features = vote_features.contiguous()
        if normalize:
            vote_features = vote_features.div(torch.norm(vote_features, p=2, dim=2).unsqueeze(2))
        return vote_xyz, vote_features
