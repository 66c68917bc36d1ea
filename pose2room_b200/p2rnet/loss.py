"""Detection losses of P2RNet on the B200 nn_distance kernel.

Same registered names, constructor `(weight, device, cfg)`, call signature and returned dict (ten 0-d tensors,
'total' differentiable) as /root/reference/models/loss.py:22-189.  The three nn_distance call sites
(loss.py:64,105,128) go through pose2room_b200.geometry.nn_distance (one kernel each, indices bit-identical
to the reference); the per-sample Python loop of compute_correspondence (loss.py:126-133, boolean-mask
indexing = a host sync per sample) is replaced by ONE batched call in which padded GT slots are moved far
away, which yields the same distances and indices because valid boxes always form a prefix
(models/p2rnet/dataloader.py:119-123).
"""
import os

import torch
from torch import nn
from torch.autograd import Function

from .. import _lib, ops
from ..geometry import huber_loss, nn_distance
from .registers import LOSSES

FAR_THRESHOLD = 0.6
NEAR_THRESHOLD = 0.3
GT_VOTE_FACTOR = 3
OBJECTNESS_CLS_WEIGHTS = [0.1, 0.9]
_FAR_AWAY = 1.0e4  # padded GT centres: (1e4)^2 * 3 = 3e8, finite in fp32, never the nearest


# The whole loss as one forward + one backward launch (csrc/loss_ops.cu): the default since round 2 (B200: parity-green on
# the reference goldens with the flag on, 9.29 -> 8.87 ms/step; profiles/r02_first_call_diag.log).  P2R_FUSED_LOSS=0 selects
# the chain of torch kernels below (kept as the second implementation the fused kernel is tested against); read at call time.
def fused_loss_enabled():
    return os.environ.get("P2R_FUSED_LOSS", "1") != "0"


def _rows(t, width):
    """(B,P,width) float32 view whose rows are contiguous -> (tensor, row stride in elements)."""
    if t.dtype != torch.float32 or t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1) or t.stride(1) < width:
        t = t.float().contiguous()
    return t, t.stride(1)


class _FusedDetectionLoss(Function):
    """est / gt tensors -> (out32 f32[8], out64 f64[2]); see include/p2r_b200.h (p2r_detection_loss) for the slots."""

    @staticmethod
    def forward(ctx, vote_xyz, center, size, heading, obj, sem, agg, skeleton, seed_inds, vote_label, vote_mask,
                gt_center, gt_mask, gt_size, gt_heading, gt_cls, origin):
        dev = center.device
        f32 = lambda t: t.detach().float().contiguous()
        vote_xyz, center, size, agg, skeleton = f32(vote_xyz), f32(center), f32(size), f32(agg), f32(skeleton)
        h64 = heading.dtype == torch.float64
        heading = heading.detach().contiguous() if h64 else f32(heading)
        obj, obj_stride = _rows(obj.detach(), 2)
        sem, sem_stride = _rows(sem.detach(), sem.shape[2])
        seed_inds = seed_inds.long().contiguous()
        vote_label, gt_center, gt_mask = f32(vote_label), f32(gt_center[:, :, 0:3]), f32(gt_mask)
        gt_size, gt_heading = f32(gt_size), f32(gt_heading)
        vote_mask, gt_cls = vote_mask.long().contiguous(), gt_cls.long().contiguous()
        b, s, j = skeleton.shape[:3]
        t, p, g, c = vote_label.shape[1], center.shape[1], gt_center.shape[1], sem.shape[2]
        out32 = torch.empty(8, dtype=torch.float32, device=dev)
        out64 = torch.empty(2, dtype=torch.float64, device=dev)
        scales = torch.empty(4, dtype=torch.float64, device=dev)
        u_vote = torch.empty(b, s, 3, dtype=torch.float32, device=dev)
        u_c1, u_c2, u_size = (torch.empty(b, p, 3, dtype=torch.float32, device=dev) for _ in range(3))
        u_head = torch.empty(b, p, 2, dtype=heading.dtype, device=dev)
        u_obj = torch.empty(b, p, 2, dtype=torch.float32, device=dev)
        u_sem = torch.empty(b, p, c, dtype=torch.float32, device=dev)
        n_ws = int(_lib.query("p2r_detection_loss_workspace", b, s))
        ws = ops.zeros_ws(n_ws, torch.float64, dev)
        with torch.cuda.device(dev):
            _lib.call("p2r_detection_loss", vote_xyz.data_ptr(), center.data_ptr(), size.data_ptr(), heading.data_ptr(),
                      int(h64), obj.data_ptr(), int(obj_stride), sem.data_ptr(), int(sem_stride), agg.data_ptr(),
                      skeleton.data_ptr(), seed_inds.data_ptr(), vote_label.data_ptr(), vote_mask.data_ptr(),
                      gt_center.data_ptr(), gt_mask.data_ptr(), gt_size.data_ptr(), gt_heading.data_ptr(),
                      gt_cls.data_ptr(), b, s, j, t, p, g, c, int(origin), out32.data_ptr(), out64.data_ptr(),
                      scales.data_ptr(), u_vote.data_ptr(), u_c1.data_ptr(), u_c2.data_ptr(), u_size.data_ptr(),
                      u_head.data_ptr(), u_obj.data_ptr(), u_sem.data_ptr(), ws.data_ptr(), n_ws,
                      torch.cuda.current_stream().cuda_stream)
        ctx.save_for_backward(scales, u_vote, u_c1, u_c2, u_size, u_head, u_obj, u_sem)
        ctx.dims = (b, s, p, c, int(h64))
        return out32, out64

    @staticmethod
    def backward(ctx, g32, g64):
        scales, u_vote, u_c1, u_c2, u_size, u_head, u_obj, u_sem = ctx.saved_tensors
        b, s, p, c, h64 = ctx.dims
        g32, g64 = g32.float().contiguous(), g64.double().contiguous()
        d_vote, d_center, d_size = torch.empty_like(u_vote), torch.empty_like(u_c1), torch.empty_like(u_size)
        d_head, d_obj, d_sem = torch.empty_like(u_head), torch.empty_like(u_obj), torch.empty_like(u_sem)
        with torch.cuda.device(g32.device):
            _lib.call("p2r_detection_loss_grad", g32.data_ptr(), g64.data_ptr(), scales.data_ptr(), u_vote.data_ptr(),
                      u_c1.data_ptr(), u_c2.data_ptr(), u_size.data_ptr(), u_head.data_ptr(), h64, u_obj.data_ptr(),
                      u_sem.data_ptr(), b, s, p, c, d_vote.data_ptr(), d_center.data_ptr(), d_size.data_ptr(),
                      d_head.data_ptr(), d_obj.data_ptr(), d_sem.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return (d_vote, d_center, d_size, d_head, d_obj, d_sem) + (None,) * 11


class BaseLoss(object):
    def __init__(self, weight=1, device=0, cfg=None):
        self.weight = weight
        self.device = device
        self.origin_joint_id = cfg.dataset_config.origin_joint_id


@LOSSES.register_module
class Null(BaseLoss):
    def __call__(self, loss):
        return self.weight * torch.mean(loss)


@LOSSES.register_module
class BoxNetDetectionLoss(BaseLoss):
    def __init__(self, weight, device, cfg=None):
        super().__init__(weight, device, cfg)
        self._obj_w = {}  # per-device copy of the class weights (made once: no H2D copy inside a captured step)
        self._sem_ce = nn.CrossEntropyLoss(reduction="none")

    def _objectness_weights(self, device):
        key = str(device)
        if key not in self._obj_w:
            self._obj_w[key] = torch.tensor(OBJECTNESS_CLS_WEIGHTS, device=device)
        return self._obj_w[key]

    def compute_vote_loss(self, est, gt):
        """loss.py:90-115: supervise each vote with the GT vote closest to any body joint of its seed."""
        b, s, j = est["seed_skeleton"].shape[:3]
        vote_xyz = est["vote_xyz"]
        seed_inds = est["seed_inds"].long()
        o = self.origin_joint_id
        mask = torch.gather(gt["vote_label_mask"][..., o], 1, seed_inds)
        votes = torch.gather(gt["vote_label"][:, :, o], 1, seed_inds.view(b, s, 1).expand(b, s, 3 * GT_VOTE_FACTOR))
        votes = est["seed_skeleton"][:, :, o:o + 1] + votes.view(b, s, GT_VOTE_FACTOR, 3)  # slice, not a list index (no H2D index copy)
        skeleton = est["seed_skeleton"].reshape(b * s, j, 3)
        _, _, dist2, ind2 = nn_distance(votes.view(b * s, GT_VOTE_FACTOR, 3), skeleton)
        pick = torch.gather(ind2, 1, dist2.argmin(-1, keepdim=True)).view(b, s, 1)
        target = torch.gather(votes, 2, pick.unsqueeze(-1).expand(b, s, 1, 3)).squeeze(2)
        vote_loss = torch.mean(huber_loss(vote_xyz - target, delta=1.0), -1)
        m = mask.float()
        return torch.sum(vote_loss * m) / (torch.sum(m) + 1e-6)

    def compute_correspondence(self, est, gt):
        """loss.py:117-150, batched."""
        agg = est["aggregated_vote_xyz"]
        gt_center = gt["center_label"][:, :, 0:3]
        box_mask = gt["box_label_mask"]
        padded = torch.where(box_mask.unsqueeze(-1) > 0, gt_center, torch.full_like(gt_center, _FAR_AWAY))
        dist1, assignment, _, _ = nn_distance(agg, padded)
        dist = torch.sqrt(dist1 + 1e-6)
        near = dist < NEAR_THRESHOLD
        objectness_label = near.long()
        objectness_mask = (near | (dist > FAR_THRESHOLD)).float()
        scores = est["objectness_scores"]
        ce = nn.functional.cross_entropy(scores.transpose(2, 1), objectness_label,
                                         weight=self._objectness_weights(scores.device), reduction="none")
        objectness_loss = torch.sum(ce * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)
        return assignment, objectness_loss, objectness_label, objectness_mask

    def compute_box_and_sem_cls_loss(self, est, gt, meta, config):
        """loss.py:42-88."""
        assignment = meta["object_assignment"]
        obj = meta["objectness_label"].float()
        denom = torch.sum(obj) + 1e-6
        box_mask = gt["box_label_mask"]
        dist1, _, dist2, _ = nn_distance(est["center"], gt["center_label"])
        center_loss = (torch.sum(dist1 * obj) / denom + torch.sum(dist2 * box_mask) / (torch.sum(box_mask) + 1e-6)) / 2.
        gt_size = torch.gather(gt["size"], 1, assignment.unsqueeze(-1).expand(-1, -1, 3))
        size_loss = torch.sum(torch.mean(huber_loss(est["size"] - gt_size, delta=1.0), -1) * obj) / denom
        gt_heading = torch.gather(gt["heading"], 1, assignment.unsqueeze(-1).expand(-1, -1, 2))
        heading_loss = torch.sum(torch.mean(huber_loss(est["heading"] - gt_heading, delta=1.0), -1) * obj) / denom
        gt_cls = torch.gather(gt["sem_cls_label"], 1, assignment)
        sem = self._sem_ce(est["sem_cls_scores"].transpose(2, 1), gt_cls)
        sem_cls_loss = torch.sum(sem * obj) / denom
        return center_loss, size_loss, heading_loss, sem_cls_loss

    def fused(self, est, gt):
        """The same ten numbers from p2r_detection_loss: 1 launch forward, 1 backward."""
        for k in ("vote_xyz", "center", "size", "heading", "objectness_scores", "sem_cls_scores"):
            if not est[k].is_cuda:
                raise RuntimeError("pose2room_b200: CUDA tensor required (there is no CPU path)")
        o32, o64 = _FusedDetectionLoss.apply(
            est["vote_xyz"], est["center"], est["size"], est["heading"], est["objectness_scores"], est["sem_cls_scores"],
            est["aggregated_vote_xyz"], est["seed_skeleton"], est["seed_inds"], gt["vote_label"], gt["vote_label_mask"],
            gt["center_label"], gt["box_label_mask"], gt["size"], gt["heading"], gt["sem_cls_label"], self.origin_joint_id)
        total, heading_loss = o64[1], o64[0]
        if est["heading"].dtype != torch.float64:    # type promotion of the torch chain: float64 only through a float64 heading
            total, heading_loss = total.float(), heading_loss.float()
        return {"total": total, "vote_loss": o32[0], "objectness_loss": o32[1], "center_loss": o32[2],
                "size_loss": o32[3], "heading_loss": heading_loss, "sem_cls_loss": o32[4], "pos_ratio": o32[5],
                "neg_ratio": o32[6], "obj_acc": o32[7]}

    def __call__(self, est, gt, dataset_config):
        if fused_loss_enabled():
            return self.fused(est, gt)

        def box_chain():
            assignment, objectness_loss, objectness_label, objectness_mask = self.compute_correspondence(est, gt)
            meta = {"object_assignment": assignment, "objectness_label": objectness_label}
            return (objectness_loss, objectness_label, objectness_mask) + \
                self.compute_box_and_sem_cls_loss(est, gt, meta, dataset_config)
        # the vote loss and the proposal losses share nothing: two chains of ~100 tiny launches, side by side when the
        # step runs multi-stream (ops.parallel_branches), sequential otherwise
        (objectness_loss, objectness_label, objectness_mask, center_loss, size_loss, heading_loss, sem_cls_loss), \
            vote_loss = ops.parallel_branches([box_chain, lambda: self.compute_vote_loss(est, gt)])
        total = 10 * vote_loss + 5 * objectness_loss + 10 * center_loss + 10 * size_loss + 10 * heading_loss + sem_cls_loss
        n_prop = float(objectness_label.shape[0] * objectness_label.shape[1])
        pos_ratio = torch.sum(objectness_label.float()) / n_prop
        neg_ratio = torch.sum(objectness_mask) / n_prop - pos_ratio
        obj_pred = torch.argmax(est["objectness_scores"], 2)
        obj_acc = torch.sum((obj_pred == objectness_label).float() * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)
        return {"total": total, "vote_loss": vote_loss, "objectness_loss": objectness_loss, "center_loss": center_loss,
                "size_loss": size_loss, "heading_loss": heading_loss, "sem_cls_loss": sem_cls_loss,
                "pos_ratio": pos_ratio, "neg_ratio": neg_ratio, "obj_acc": obj_acc}
