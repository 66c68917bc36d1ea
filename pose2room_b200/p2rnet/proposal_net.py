"""ProposalNet: vote aggregation (one PointNet++ set-abstraction layer) + probabilistic box heads.

Module surface / state-dict of /root/reference/models/p2rnet/modules/proposal_net.py:36-252:
vote_aggregation = PointnetSAModuleVotes(npoint=num_target, radius=0.3, nsample=16, mlp=[256,256,256],
use_xyz=False, normalize_xyz=True, bn=False); conv_{center,heading,size} = 2 x SingleConv 'cbr';
conv_sem_obj = 2 x 'cbr' + 'c' -> 2 + num_class; gmm_{center,size,heading} = CategoryEmbeddingMDN.

B200 re-design of the SA layer: FPS / ball-query through the index-exact kernels, then the features are
grouped CHANNEL-LAST ([B,128,16,256] rows of 256 contiguous values, 128-bit copies) so the shared MLP is two
row-major GEMMs with the ReLU fused in the epilogue and the max over nsample is a strided column max with a
saved arg-max for the backward -- the (B,256,128,16) tensor of the reference is never laid out channel-first.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops, pointnet2_utils
from ..pointnet2_modules import PointnetSAModuleVotes
from .mdn import CategoryEmbeddingMDN, Struct
from .registers import MODULES
from .sub_modules import SingleConv, run_rows


def farthest_point_sample_torch(xyz, npoint):
    """Init-time helper (random start), as net_utils/libs.py:152-176: used once to thin the GMM mu grids."""
    b, n, _ = xyz.shape
    centroids = torch.zeros(b, npoint, dtype=torch.long)
    distance = torch.ones(b, n, dtype=xyz.dtype) * 1e10
    farthest = torch.randint(0, n, (b,), dtype=torch.long)
    batch = torch.arange(b, dtype=torch.long)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[batch, farthest, :].view(b, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = torch.max(distance, -1)[1]
    return centroids


def decode_scores(pred_center, pred_size, pred_heading, sem_obj, end_points):
    """proposal_net.py:15-34 on row-major heads: pred_* (B,P,d), sem_obj (B,P,2+C)."""
    end_points["center"] = end_points["aggregated_vote_xyz"] + pred_center
    end_points["size"] = pred_size
    end_points["heading"] = pred_heading
    end_points["objectness_scores"] = sem_obj[..., 0:2]
    end_points["sem_cls_scores"] = sem_obj[..., 2:]
    return end_points


@MODULES.register_module
class ProposalNet(nn.Module):
    def __init__(self, cfg, optim_spec=None):
        super().__init__()
        self.optim_spec = optim_spec
        self.cfg = cfg
        self.num_class = cfg.dataset_config.num_class
        self.num_proposals = cfg.config["data"]["num_target"]
        self.sampling = cfg.config["data"]["cluster_sampling"]
        self.precision = cfg.config.get("precision", "fp32")
        vote_dim = 256
        if cfg.config["mode"] != "train":
            self.multi_mode = cfg.eval_config["multi_mode"]
            self.n_samples = np.random.choice(np.arange(1, 100), 1)[0]   # proposal_net.py:57-59
        self.vote_aggregation = PointnetSAModuleVotes(npoint=self.num_proposals, radius=0.3, nsample=16,
                                                      mlp=[256, 256, vote_dim], use_xyz=False, normalize_xyz=True,
                                                      bn=False)
        sem_obj_dim = 2 + self.num_class
        gmm_dim = 128

        def head():
            return nn.Sequential(SingleConv(vote_dim, 128, order="cbr"), SingleConv(128, gmm_dim, order="cbr"))
        self.conv_center = head()
        self.conv_heading = head()
        self.conv_size = head()
        self.conv_sem_obj = nn.Sequential(SingleConv(vote_dim, 128, order="cbr"), SingleConv(128, 128, order="cbr"),
                                          SingleConv(128, sem_obj_dim, order="c"))
        ng = cfg.config["data"]["num_gaussian"]
        self.gmm_center = self.load_gmm(ng, gmm_dim, 3, "center")
        self.gmm_size = self.load_gmm(ng, gmm_dim, 3, "size")
        self.gmm_heading = self.load_gmm(ng, gmm_dim, 2, "heading")

    # ---- GMM initialisation (proposal_net.py:96-148), same grids / same thinning -------------------
    def init_mu(self, num_gaussian, kind):
        if kind == "center":
            n_theta = np.ceil(np.sqrt(num_gaussian / 2)).astype(np.uint16)
            n_phi = 2 * n_theta
            width = np.pi / n_theta
            phi = [width * i - np.pi for i in range(0, n_phi)]
            theta = np.linspace(0, np.pi, n_theta + 2)[1:-1]
            grid = np.array(np.meshgrid(phi, theta)).reshape(2, -1).T
            pts = np.hstack([0.1 * np.sin(grid[:, [1]]) * np.cos(grid[:, [0]]),
                             0.1 * np.sin(grid[:, [1]]) * np.sin(grid[:, [0]]),
                             0.1 * np.cos(grid[:, [1]])])
            mu = torch.from_numpy(pts)
            if num_gaussian < mu.size(0):
                mu = self.get_farthest_points(mu, num_gaussian)
            return mu
        if kind == "size":
            per_dim = np.ceil(num_gaussian ** (1 / 3)).astype(np.uint32)
            g = np.linspace(0.05, 3, per_dim)
            grid = np.log(np.array(np.meshgrid(g, g, g)).reshape(3, -1).T)
            return self.get_farthest_points(torch.from_numpy(grid), num_gaussian)
        if kind == "heading":
            width = 2 * np.pi / num_gaussian
            th = [width * i - np.pi for i in range(0, num_gaussian)]
            return torch.from_numpy(np.array([[np.sin(t), np.cos(t)] for t in th]))   # float64, like the reference
        raise ValueError(kind)

    @staticmethod
    def get_farthest_points(xyz, npoint):
        xyz = xyz.unsqueeze(0).float() if xyz.dim() == 2 else xyz.float()
        inds = torch.sort(farthest_point_sample_torch(xyz, npoint), dim=-1)[0]
        out = xyz[torch.arange(xyz.size(0))[:, None], inds]
        return out.squeeze(0) if out.size(0) == 1 else out

    def load_gmm(self, num_gaussian, in_dim, out_dim, kind):
        mdn_config = Struct(num_gaussian=num_gaussian, out_dim=out_dim, mu_bias_init=self.init_mu(num_gaussian, kind),
                            n_samples=1, central_tendency="mean")
        config = Struct(embedding_dims=[], out_dim=3, continuous_dim=in_dim, batch_norm_continuous_input=False,
                        hidden_dim=128, mdn_config=mdn_config)
        return CategoryEmbeddingMDN(config)

    # ---- vote aggregation on the B200 kernels ------------------------------------------------------
    def _aggregate(self, xyz, feats_rows, inds=None):
        """xyz (B,N,3) f32, feats_rows (B,N,C) channel-last -> new_xyz (B,P,3), new feats (B,P,C'), inds (B,P) i32."""
        sa = self.vote_aggregation
        b, n, c = feats_rows.shape
        xyz = xyz.contiguous()
        if inds is None:
            inds = pointnet2_utils.furthest_point_sample(xyz, sa.npoint)
        new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        idx = pointnet2_utils.ball_query(sa.radius, sa.nsample, xyz, new_xyz)          # (B,P,S) i32
        assert not sa.use_xyz and sa.pooling == "max"
        act = torch.bfloat16 if self.precision == "bf16" else torch.float32
        convs = [m for m in sa.mlp_module if isinstance(m, nn.Conv2d)]
        rows = feats_rows.to(act)
        if ops.sa_fused_available(rows, idx, convs):      # gather -> MLP -> max as ONE tcgen05 kernel (csrc/sa_fused.cu)
            pooled = ops.sa_fused(rows, idx, convs[0], convs[1])
            return new_xyz, pooled.float().reshape(b, sa.npoint, -1), inds
        grouped = ops.group_rows(rows, idx)                                            # (B,P,S,C)
        h = grouped.reshape(b * sa.npoint * sa.nsample, c)
        for conv in convs:                                                             # Conv2d 1x1 + bias + ReLU
            h = ops.linear(h, conv.weight.reshape(conv.out_channels, conv.in_channels), conv.bias, relu=True)
        pooled = ops.maxpool_rows(h.reshape(b * sa.npoint, sa.nsample, h.shape[1]))    # (B*P, C')
        return new_xyz, pooled.float().reshape(b, sa.npoint, -1), inds

    def _heads(self, xyz, features, end_points, generate):
        b = xyz.shape[0]
        if self.sampling == "vote_fps":
            xyz, feats, fps_inds = self._aggregate(xyz, features)
            sample_inds, arg = torch.sort(fps_inds, dim=-1)
            arg = arg.long()
            xyz = torch.gather(xyz, 1, arg.unsqueeze(-1).expand(-1, -1, 3))
            feats = torch.gather(feats, 1, arg.unsqueeze(-1).expand(-1, -1, feats.size(2)))
        elif self.sampling == "seed_fps":
            seed_xyz = end_points["seed_xyz"]
            move = torch.norm(torch.diff(seed_xyz, dim=1), dim=2)
            cum = torch.cumsum(torch.cat([torch.zeros(b, 1, device=xyz.device), move], dim=1), dim=1)
            step = cum[:, -1] / (self.num_proposals - 1)
            target = step.unsqueeze(-1) * torch.arange(self.num_proposals, dtype=torch.float, device=xyz.device)
            sample_inds = torch.argmin(torch.abs(cum.unsqueeze(-1) - target.unsqueeze(1)), dim=1).type(torch.int32)
            xyz, feats, _ = self._aggregate(xyz, features, sample_inds)
        else:
            raise NotImplementedError("Undefined sampling strategy.")
        end_points["aggregated_vote_xyz"] = xyz
        end_points["aggregated_vote_inds"] = sample_inds.type(torch.int64)

        p = feats.shape[1]
        act = torch.bfloat16 if self.precision == "bf16" else torch.float32
        rows = feats.reshape(b * p, -1).to(act)
        # four independent chains (each ~30 small launches, again in the backward): side by side when the step runs
        # multi-stream (ops.overlap_weight_grads), one after the other otherwise.  The train-mode noise is drawn in
        # the reference's order (center, size, heading: mdn.py:44) because the host issues the chains in that order.
        def gmm(head, conv):
            def run():
                f = run_rows(conv, rows)
                if generate:
                    return head.generate_rows(f, self.multi_mode, self.n_samples)
                return head.predict_rows(f), None
            return run
        (c, pi_c), (s, pi_s), (h, pi_h), sem_obj = ops.parallel_branches([
            gmm(self.gmm_center, self.conv_center), gmm(self.gmm_size, self.conv_size),
            gmm(self.gmm_heading, self.conv_heading),
            lambda: run_rows(self.conv_sem_obj, rows).float().reshape(b, p, -1)])
        end_points = decode_scores(c.reshape(b, p, 3), s.reshape(b, p, 3), h.reshape(b, p, 2), sem_obj, end_points)
        if generate:
            # reference layout of pi: (B, G, P)  (proposal_net.py:239-247)
            end_points["pi"] = {"center": pi_c.reshape(b, p, -1).transpose(1, 2),
                                "size": pi_s.reshape(b, p, -1).transpose(1, 2),
                                "heading": pi_h.reshape(b, p, -1).transpose(1, 2)}
        return end_points, feats

    def forward(self, xyz, features, end_points, export_proposal_feature=False):
        """xyz (B,K,3), features (B,K,C) -> end_points (+ proposal features (B,P,C) on request)."""
        end_points, feats = self._heads(xyz, features, end_points, generate=False)
        return end_points, (feats.contiguous() if export_proposal_feature else None)

    def generate(self, xyz, features, end_points, export_proposal_feature=False):
        end_points, feats = self._heads(xyz, features, end_points, generate=True)
        return end_points, (feats.contiguous() if export_proposal_feature else None)
