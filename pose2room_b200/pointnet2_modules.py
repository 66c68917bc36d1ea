"""Set-abstraction / feature-propagation modules of the reference's `pointnet2_modules`, on the B200 ops.

Mirrors /root/reference/external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:
  build_shared_mlp       :9-19     (Conv2d 1x1 [+BN] + ReLU stack; state-dict keys mlp_module.{0,2,..})
  PointnetSAModuleVotes  :150-261  (FPS -> gather -> ball query + group -> shared MLP -> pool)
  PointnetFPModule       :346-406  (three_nn -> inverse-distance weights -> three_interpolate -> MLP)
Constructor arguments, forward signatures, return tuples and parameter names are the reference's, so a
reference state-dict loads unchanged.  (MSG / STN variants of the reference file have no caller in the
P2RNet path -- SURVEY.md section 2.3 -- and are not rebuilt.)
"""
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, pointnet2_utils


def build_shared_mlp(mlp_spec: List[int], bn: bool = True):
    layers = []
    for i in range(1, len(mlp_spec)):
        layers.append(nn.Conv2d(mlp_spec[i - 1], mlp_spec[i], kernel_size=1, bias=not bn))
        if bn:
            layers.append(nn.BatchNorm2d(mlp_spec[i]))
        layers.append(nn.ReLU(True))
    return nn.Sequential(*layers)


def _bn_rows_supported(c):
    return 256 % c == 0 or c % 256 == 0           # channel counts the column-statistics kernels take (csrc/dense_ops.cu)


def run_shared_mlp(mlp_module, x, pool_max=False):
    """A build_shared_mlp stack applied to x (B, C, P, S) on this library's kernels: every 1x1 Conv2d is a row-major GEMM
    (ops.linear: tcgen05 for bf16 rows, the fixed-order SIMT kernel for fp32), BatchNorm2d a column statistic
    (ops.batchnorm_act), ReLU fused into whichever comes last; pool_max also takes the max over S (ops.maxpool_rows).
    Returns (B, C', P, S), or (B, C', P, 1) with pool_max.  Falls back to the nn modules (cuDNN) only for inputs that are
    not on a GPU-supported shape (BatchNorm widths the statistics kernels do not take)."""
    mods = list(mlp_module)
    convs = [m for m in mods if isinstance(m, nn.Conv2d)]
    ok = x.is_cuda and x.dtype == torch.float32 and all(
        isinstance(m, (nn.Conv2d, nn.BatchNorm2d, nn.ReLU)) for m in mods) and all(
        m.kernel_size == (1, 1) for m in convs) and all(
        _bn_rows_supported(m.num_features) for m in mods if isinstance(m, nn.BatchNorm2d))
    if not ok:
        y = mlp_module(x)
        return F.max_pool2d(y, kernel_size=[1, y.size(3)]) if pool_max else y
    b, c, p, s = x.shape
    rows = x.permute(0, 2, 3, 1).reshape(b * p * s, c)
    i = 0
    while i < len(mods):
        conv = mods[i]
        i += 1
        bn = relu = None
        if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
            bn = mods[i]
            i += 1
        if i < len(mods) and isinstance(mods[i], nn.ReLU):
            relu = True
            i += 1
        w = conv.weight.reshape(conv.out_channels, conv.in_channels)
        if bn is None:
            rows = ops.linear(rows, w, conv.bias, relu=bool(relu))
        else:
            rows = ops.batchnorm_act(ops.linear(rows, w, conv.bias), bn, relu=bool(relu))
    if pool_max:
        return ops.maxpool_rows(rows.reshape(b * p, s, rows.shape[1])).reshape(b, p, 1, -1).permute(0, 3, 1, 2)
    return rows.reshape(b, p, s, -1).permute(0, 3, 1, 2)


class PointnetSAModuleVotes(nn.Module):
    """Set abstraction that also returns the sampled indices (pointnet2_modules.py:150-261)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = "max", sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.pooling = pooling
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else (radius / 2 if radius is not None else None)
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                                                         normalize_xyz=normalize_xyz,
                                                         sample_uniformly=sample_uniformly,
                                                         ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = list(mlp)
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3
        self.mlp_module = build_shared_mlp(mlp_spec, bn=bn)

    def forward(self, xyz, features=None, inds=None):
        """xyz (B,N,3), features (B,C,N) [, inds (B,npoint) i32] ->
        new_xyz (B,npoint,3), new_features (B,mlp[-1],npoint), inds (B,npoint) [, unique_cnt]."""
        xyz_flipped = xyz.transpose(1, 2).contiguous()
        if inds is None:
            inds = pointnet2_utils.furthest_point_sample(xyz, self.npoint) if self.npoint is not None else None
        else:
            assert inds.shape[1] == self.npoint
        new_xyz = (pointnet2_utils.gather_operation(xyz_flipped, inds).transpose(1, 2).contiguous()
                   if self.npoint is not None else None)
        unique_cnt = None
        if self.ret_unique_cnt:
            grouped_features, grouped_xyz, unique_cnt = self.grouper(xyz, new_xyz, features)
        else:
            grouped_features, grouped_xyz = self.grouper(xyz, new_xyz, features)
        if self.pooling == "max":    # shared MLP + max over nsample on this library's GEMM / pooling kernels
            new_features = run_shared_mlp(self.mlp_module, grouped_features, pool_max=True)
        else:
            new_features = run_shared_mlp(self.mlp_module, grouped_features)  # (B, mlp[-1], npoint, nsample)
        if self.pooling == "max":
            pass
        elif self.pooling == "avg":
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
        elif self.pooling == "rbf":
            rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (self.sigma ** 2) / 2)
            new_features = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(self.nsample)
        new_features = new_features.squeeze(-1)
        if self.ret_unique_cnt:
            return new_xyz, new_features, inds, unique_cnt
        return new_xyz, new_features, inds


class PointnetFPModule(nn.Module):
    """Feature propagation (pointnet2_modules.py:346-406)."""

    def __init__(self, mlp, bn=True):
        super().__init__()
        self.mlp = build_shared_mlp(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        new_features = torch.cat([interpolated, unknow_feats], dim=1) if unknow_feats is not None else interpolated
        return run_shared_mlp(self.mlp, new_features.unsqueeze(-1)).squeeze(-1)
