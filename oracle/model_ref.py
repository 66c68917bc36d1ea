"""Plain fp32 PyTorch (CPU) restatement of the reference's P2RNet forward / generate / loss.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by pose2room_b200/.  It exists because /root/reference cannot travel to the
GPU box: this file is the "port" whose outputs are pinned, in the build container, against the real
reference (tests/test_model_oracle.py, tests/golden/p2rnet_*.npz) and which then referees the CUDA path.

It follows the reference's own computational structure and tensor layouts (channel-first Conv1d/Conv2d,
1x1 conv 64 -> 704 followed by einsum 'nkctv,kvw->nctw', materialised (B,256,128,16) grouped tensor,
repeat-based nn_distance, per-sample correspondence loop) so that timing it on host cores is a fair CPU
baseline of the reference algorithm.  Parameters are addressed by the reference's state-dict keys.

Citations (into /root/reference): models/p2rnet/modules/stgcn.py:84-152, stgcn_layers.py:58-67,433-439,
vote_center.py:34-59, proposal_net.py:150-198, mdn.py:34-99,112-125, network.py:44-106, models/loss.py:42-189,
external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:220-256, pointnet2_utils.py:319-346.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import geometry_ref
from .pointnet2_ref import RefExt


class _RefGrouping(torch.autograd.Function):
    """pointnet2_utils.py:194-240 (GroupingOperation) over an `_ext`-shaped module holding the reference's kernels."""

    @staticmethod
    def forward(ctx, ext, features, idx):
        ctx.ext, ctx.n = ext, features.size(2)
        ctx.save_for_backward(idx)
        return ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return None, ctx.ext.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


class RefP2RNet:
    def __init__(self, state_dict, joint_num, num_seeds=512, num_target=128, training=True, origin_joint_id=0,
                 num_class=22, bn_momentum=0.1, device="cpu", ext=None):
        """device / ext: the default is the CPU port (oracle/pointnet2_ref.c behind RefExt).  bench.py's `gpu_reference`
        leg passes device="cuda" and the UNMODIFIED reference extension compiled for sm_100a (oracle/_ref/p2r_ref_ext.so,
        oracle/build_ref_ext.py): the reference's own kernels + the same fp32 cuDNN / cuBLAS calls the reference makes --
        "the reference on B200" (second baseline)."""
        self.dev = torch.device(device)
        self.ext = ext if ext is not None else RefExt
        self.p = {}
        for k, v in state_dict.items():
            t = v.detach().clone().to(self.dev)
            if t.is_floating_point() and "running_" not in k and k != "backbone.A":
                t.requires_grad_(True)
            self.p[k] = t
        self.J, self.S, self.P = joint_num, num_seeds, num_target
        self.training = training
        self.o = origin_joint_id
        self.num_class = num_class
        self.mom = bn_momentum

    def parameters(self):
        return [t for t in self.p.values() if t.requires_grad]

    def zero_grad(self):
        for t in self.parameters():
            t.grad = None

    # ------------------------------------------------------------------ building blocks
    def _bn(self, x, prefix):
        return F.batch_norm(x, self.p[prefix + ".running_mean"], self.p[prefix + ".running_var"],
                            self.p[prefix + ".weight"], self.p[prefix + ".bias"], self.training, self.mom, 1e-5)

    def _single(self, x, prefix, order):  # SingleConv on (B,C,L)
        x = F.conv1d(x, self.p[prefix + ".conv.weight"], self.p.get(prefix + ".conv.bias"))
        if "b" in order:
            x = self._bn(x, prefix + ".batchnorm")
        if "r" in order:
            x = F.relu(x)
        return x

    def _mlp3(self, x, prefix):
        x = self._single(x, prefix + ".0", "cbr")
        x = self._single(x, prefix + ".1", "cbr")
        return self._single(x, prefix + ".2", "c")

    # ------------------------------------------------------------------ backbone (stgcn.py:84-152)
    def backbone(self, joints):
        B, T, J, D = joints.shape
        hip = joints[:, :, self.o]
        if self.S >= T:
            seed_inds = torch.round(torch.linspace(0, T - 1, self.S)).long().repeat(B, 1).to(self.dev)
        else:
            move = torch.norm(torch.diff(hip, dim=1), dim=2)
            cum = torch.cumsum(torch.cat([torch.zeros(B, 1, device=self.dev), move], dim=1), dim=1)
            step = cum[:, -1] / (self.S - 1)
            target = step.unsqueeze(-1) * torch.arange(self.S, dtype=torch.float, device=self.dev)
            seed_inds = torch.argmin(torch.abs(cum.unsqueeze(-1) - target.unsqueeze(1)), dim=1)
        x = joints - joints[:, :, [self.o]]
        k = 20
        idx = (torch.arange(T)[:, None] + torch.arange(-k // 2, k // 2)[None]).clamp(0, T - 1).to(self.dev)   # (T,k); built on the host each call like stgcn.py:109-114
        rel = hip[:, idx] - hip[:, :, None]                                                        # (B,T,k,3)
        pos = self._mlp3(rel.reshape(B, T * k, 3).transpose(1, 2), "backbone.pos_embed")
        pos = pos.transpose(1, 2).contiguous().view(B, T, k, -1).mean(dim=2)
        sk = self._mlp3(x.reshape(B, T * J, 3).transpose(1, 2), "backbone.sk_feat")
        sk = sk.transpose(1, 2).contiguous().view(B, T, J, -1)
        x = (sk + pos.unsqueeze(2)).permute(0, 3, 1, 2).contiguous()                               # (B,64,T,J)
        A = self.p["backbone.A"]
        for i in range(6):
            pre = "backbone.st_gcn_networks.%d" % i
            Ai = A * self.p["backbone.edge_importance.%d" % i]
            res = x if i > 0 else 0
            y = F.conv2d(x, self.p[pre + ".gcn.conv.weight"], self.p[pre + ".gcn.conv.bias"])
            n, kc, t, v = y.shape
            y = torch.einsum("nkctv,kvw->nctw", y.view(n, A.shape[0], kc // A.shape[0], t, v), Ai).contiguous()
            y = F.relu(self._bn(y, pre + ".tcn.0"))
            y = F.conv2d(y, self.p[pre + ".tcn.2.weight"], self.p[pre + ".tcn.2.bias"], padding=(1, 0))
            y = self._bn(y, pre + ".tcn.3")
            x = F.relu(y + res)
        x = x.transpose(2, 3).contiguous().view(B, x.size(1) * J, T)
        x = F.conv1d(x, self.p["backbone.conv_joint.weight"], self.p["backbone.conv_joint.bias"]).transpose(1, 2)
        seed_skeleton = torch.gather(joints, 1, seed_inds[:, :, None, None].expand(B, self.S, J, D))
        seed_features = torch.gather(x, 1, seed_inds[:, :, None].expand(B, self.S, x.size(-1)))
        return {"seed_inds": seed_inds, "seed_skeleton": seed_skeleton[..., :3], "seed_features": seed_features}

    # ------------------------------------------------------------------ voting (vote_center.py:34-59)
    def centervoting(self, seed_skeleton, seed_features):
        seed_xyz = seed_skeleton[:, :, self.o]
        B, S, _ = seed_xyz.shape
        net = seed_features.transpose(1, 2)
        net = self._single(net, "centervoting.conv_input.0", "cbr")
        net = self._single(net, "centervoting.conv_input.1", "cbr")
        net = self._single(net, "centervoting.conv_input.2", "c").transpose(2, 1)
        vote_xyz = seed_xyz + net[:, :, 0:3]
        vote_features = seed_features + net[:, :, 3:]
        return vote_xyz.contiguous(), vote_features.contiguous()

    # ------------------------------------------------------------------ detection (proposal_net.py:150-252)
    def _sa(self, xyz, features):  # features (B,C,N)
        B, C, N = features.shape
        inds = self.ext.furthest_point_sampling(xyz.detach().contiguous(), self.P)
        new_xyz = torch.gather(xyz, 1, inds.long()[:, :, None].expand(B, self.P, 3)).contiguous()
        idx = self.ext.ball_query(new_xyz.detach().contiguous(), xyz.detach().contiguous(), 0.3, 16)       # (B,P,S) i32
        if self.dev.type == "cuda":       # the reference's GroupingOperation (pointnet2_utils.py:194-240) on its own kernels
            grouped = _RefGrouping.apply(self.ext, features.contiguous(), idx.contiguous())
        else:
            grouped = torch.gather(features[:, :, None, :].expand(B, C, self.P, N), 3,
                                   idx.long()[:, None].expand(B, C, self.P, 16))                     # (B,C,P,16)
        pre = "detection.vote_aggregation.mlp_module"
        h = F.relu(F.conv2d(grouped, self.p[pre + ".0.weight"], self.p[pre + ".0.bias"]))
        h = F.relu(F.conv2d(h, self.p[pre + ".2.weight"], self.p[pre + ".2.bias"]))
        h = F.max_pool2d(h, kernel_size=[1, h.size(3)]).squeeze(-1)                                  # (B,C,P)
        return new_xyz, h, inds

    def _head(self, x, prefix, n):
        for i in range(n - 1):
            x = self._single(x, "%s.%d" % (prefix, i), "cbr")
        return x

    def _gmm(self, feat, name, generate):
        pre = "detection.%s" % name
        h = self._single(feat, pre + ".backbone", "cbr")
        pi = torch.sigmoid(self._single(h, pre + ".mdn.pi", "c"))                                    # (B,G,P)
        B, G, P = pi.shape
        pi_r = pi.transpose(1, 2).contiguous().view(B * P, G)
        mu = self.p[pre + ".mdn.mu"]
        if generate:
            out = torch.sum(mu[None] * pi_r[:, :, None], dim=1)
        else:
            sigma = torch.exp(self.p[pre + ".mdn.log_sigma"])[None, :, None, :].expand(B * P, G, 1, -1)
            mu_e = mu[None, :, None, :].expand(B * P, G, 1, -1)
            eps = mu_e.data.new(mu_e.size()).normal_()
            out = torch.sum((eps * sigma + mu_e) * pi_r[:, :, None, None], dim=1).mean(dim=1)
        return out.view(B, P, -1), pi

    def detection(self, xyz, features, end_points, generate=False):
        features = features.transpose(1, 2).contiguous()
        xyz, features, fps_inds = self._sa(xyz, features)
        sample_inds, arg = torch.sort(fps_inds, dim=-1)
        arg = arg.long()
        xyz = torch.gather(xyz, 1, arg.unsqueeze(-1).repeat(1, 1, 3))
        features = torch.gather(features, 2, arg.unsqueeze(1).repeat(1, features.size(1), 1))
        end_points["aggregated_vote_xyz"] = xyz
        end_points["aggregated_vote_inds"] = sample_inds.type(torch.int64)
        cf = self._head(features, "detection.conv_center", 3)
        sf = self._head(features, "detection.conv_size", 3)
        hf = self._head(features, "detection.conv_heading", 3)
        so = self._head(features, "detection.conv_sem_obj", 3)
        so = self._single(so, "detection.conv_sem_obj.2", "c").transpose(2, 1)
        c, pc = self._gmm(cf, "gmm_center", generate)
        s, ps = self._gmm(sf, "gmm_size", generate)
        h, ph = self._gmm(hf, "gmm_heading", generate)
        end_points["center"] = xyz + c
        end_points["size"] = s
        end_points["heading"] = h
        end_points["objectness_scores"] = so[..., 0:2]
        end_points["sem_cls_scores"] = so[..., 2:]
        if generate:
            end_points["pi"] = {"center": pc, "size": ps, "heading": ph}
        return end_points

    # ------------------------------------------------------------------ model API (network.py:44-106)
    def forward(self, data, generate=False):
        ep = self.backbone(data["input_joints"])
        xyz, feats = self.centervoting(ep["seed_skeleton"], ep["seed_features"])
        feats = feats.div(torch.norm(feats, p=2, dim=2).unsqueeze(2))
        ep["vote_xyz"], ep["vote_features"] = xyz, feats
        return self.detection(xyz, feats, ep, generate)

    def generate(self, data):
        with torch.no_grad():
            ep = self.forward(data, generate=True)
        hip = data["input_joints"][:, :, self.o].cpu().numpy()
        parsed = geometry_ref.parse_predictions(ep["center"].cpu().numpy(), ep["size"].cpu().numpy(),
                                                ep["heading"].cpu().numpy(), ep["objectness_scores"].cpu().numpy(),
                                                ep["sem_cls_scores"].cpu().numpy(), hip)
        return ep, parsed

    # ------------------------------------------------------------------ loss (models/loss.py:42-189)
    @staticmethod
    def _nn_distance(pc1, pc2):
        N, M = pc1.shape[1], pc2.shape[1]
        diff = pc1.unsqueeze(2).repeat(1, 1, M, 1) - pc2.unsqueeze(1).repeat(1, N, 1, 1)
        dist = torch.sum(diff ** 2, dim=-1)
        d1, i1 = torch.min(dist, dim=2)
        d2, i2 = torch.min(dist, dim=1)
        return d1, i1, d2, i2

    @staticmethod
    def _huber(err, delta=1.0):
        a = torch.abs(err)
        q = torch.clamp(a, max=delta)
        return 0.5 * q ** 2 + delta * (a - q)

    def loss(self, est, gt):
        B, S, J = est["seed_skeleton"].shape[:3]
        o = self.o
        # vote loss
        seed_inds = est["seed_inds"].long()
        vmask = torch.gather(gt["vote_label_mask"][..., o], 1, seed_inds)
        votes = torch.gather(gt["vote_label"][:, :, o], 1, seed_inds.view(B, S, 1).repeat(1, 1, 9)).view(B, S, 3, 3)
        votes = est["seed_skeleton"][:, :, [o]] + votes
        _, _, d2, i2 = self._nn_distance(votes.view(B * S, 3, 3), est["seed_skeleton"].reshape(B * S, J, 3))
        pick = torch.gather(i2, 1, d2.argmin(-1).unsqueeze(-1)).view(B, S, 1)
        target = torch.gather(votes, 2, pick.unsqueeze(-1).repeat(1, 1, 1, 3)).squeeze(2)
        vote_loss = torch.mean(self._huber(est["vote_xyz"] - target), -1)
        vote_loss = torch.sum(vote_loss * vmask.float()) / (torch.sum(vmask.float()) + 1e-6)
        # correspondence (per-sample loop, like the reference)
        d1s, assigns = [], []
        for xyz_b, c_b, m_b in zip(est["aggregated_vote_xyz"], gt["center_label"][:, :, 0:3], gt["box_label_mask"]):
            d1, i1, _, _ = self._nn_distance(xyz_b.unsqueeze(0), c_b[m_b > 0].unsqueeze(0))
            d1s.append(d1)
            assigns.append(i1)
        dist1 = torch.cat(d1s, 0)
        assignment = torch.cat(assigns, 0)
        eu = torch.sqrt(dist1 + 1e-6)
        obj_label = (eu < 0.3).long()
        obj_mask = ((eu < 0.3) | (eu > 0.6)).float()
        ce = F.cross_entropy(est["objectness_scores"].transpose(2, 1), obj_label, weight=torch.tensor([0.1, 0.9], device=est["objectness_scores"].device),
                             reduction="none")
        objectness_loss = torch.sum(ce * obj_mask) / (torch.sum(obj_mask) + 1e-6)
        # box + class
        objf = obj_label.float()
        den = torch.sum(objf) + 1e-6
        bm = gt["box_label_mask"]
        d1, _, d2, _ = self._nn_distance(est["center"], gt["center_label"])
        center_loss = (torch.sum(d1 * objf) / den + torch.sum(d2 * bm) / (torch.sum(bm) + 1e-6)) / 2.
        gsize = torch.gather(gt["size"], 1, assignment.unsqueeze(-1).repeat(1, 1, 3))
        size_loss = torch.sum(torch.mean(self._huber(est["size"] - gsize), -1) * objf) / den
        ghead = torch.gather(gt["heading"], 1, assignment.unsqueeze(-1).repeat(1, 1, 2))
        heading_loss = torch.sum(torch.mean(self._huber(est["heading"] - ghead), -1) * objf) / den
        gcls = torch.gather(gt["sem_cls_label"], 1, assignment)
        sem = F.cross_entropy(est["sem_cls_scores"].transpose(2, 1), gcls, reduction="none")
        sem_cls_loss = torch.sum(sem * objf) / den
        total = 10 * vote_loss + 5 * objectness_loss + 10 * center_loss + 10 * size_loss + 10 * heading_loss + sem_cls_loss
        n = float(obj_label.numel())
        pos_ratio = torch.sum(objf) / n
        neg_ratio = torch.sum(obj_mask) / n - pos_ratio
        acc = torch.sum((torch.argmax(est["objectness_scores"], 2) == obj_label).float() * obj_mask) / (torch.sum(obj_mask) + 1e-6)
        return {"total": total, "vote_loss": vote_loss, "objectness_loss": objectness_loss, "center_loss": center_loss,
                "size_loss": size_loss, "heading_loss": heading_loss, "sem_cls_loss": sem_cls_loss,
                "pos_ratio": pos_ratio, "neg_ratio": neg_ratio, "obj_acc": acc}
