"""Seeded synthetic pose sequences + ground truth in the reference's `data` dict schema.

Host-side data plumbing (numpy, no kernels).  Schema = what `P2RNet_VirtualHome.__getitem__`
collates (/root/reference/models/p2rnet/dataloader.py:137-146): input_joints f32 (B,T,J,3),
box_label_mask f32 (B,10), sem_cls_label i64 (B,10), center_label f32 (B,10,3), size f32 (B,10,3)
= log size, heading f32 (B,10,2) = (sin, cos), vote_label f32 (B,T,J,9), vote_label_mask i64 (B,T,J).

Generator follows SURVEY.md section 8(d): hip = cumulative N(0, 0.05^2) steps with y clamped to
[0.8, 1.0]; other joints = hip + N(0, 0.3^2), joint 0 = hip; 1..10 GT boxes per scene with centres
= hip at random frames + N(0, 0.3^2), size ~ U(0.2, 1.7), heading ~ U(-pi, pi), class ~ U{0..21};
votes = (centre - joint) of the first <=3 boxes whose OBB enlarged by 1.0 m contains the joint, the
first vote replicated into unused slots (schema of utils/virtualhome/3_generate_samples.py:56-79).
"""
import numpy as np

MAX_GT = 10
NUM_CLASS = 22
CONTACT = 1.0


def _rot(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]])


def make_scene(rng, T, J, raw=False):
    steps = rng.normal(0.0, 0.05, size=(T, 3))
    hip = np.cumsum(steps, axis=0)
    hip[:, 1] = np.clip(0.9 + hip[:, 1], 0.8, 1.0)
    joints = hip[:, None, :] + rng.normal(0.0, 0.3, size=(T, J, 3))
    joints[:, 0] = hip

    n_gt = int(rng.integers(1, MAX_GT + 1))
    frames = rng.integers(0, T, size=n_gt)
    centers = hip[frames] + rng.normal(0.0, 0.3, size=(n_gt, 3))
    sizes = rng.uniform(0.2, 1.7, size=(n_gt, 3))
    theta = rng.uniform(-np.pi, np.pi, size=n_gt)
    classes = rng.integers(0, NUM_CLASS, size=n_gt)

    votes = np.zeros((T, J, 9))
    mask = np.zeros((T, J), np.int64)
    slot = np.zeros((T, J), np.int64)
    flat = joints.reshape(-1, 3)
    for g in range(n_gt):
        local = (flat - centers[g]) @ _rot(theta[g]).T
        inside = np.all(np.abs(local) <= sizes[g] / 2.0 + CONTACT, axis=1).reshape(T, J)
        v = centers[g][None, None, :] - joints
        tt, jj = np.nonzero(inside)
        for t, j in zip(tt, jj):
            s = slot[t, j]
            votes[t, j, 3 * s:3 * s + 3] = v[t, j]
            if s == 0:
                votes[t, j, 3:6] = v[t, j]
                votes[t, j, 6:9] = v[t, j]
        mask[inside] = 1
        slot[inside] = np.minimum(2, slot[inside] + 1)

    if raw:
        return dict(joints=joints, votes=votes, mask=mask, centers=centers, sizes=sizes, theta=theta,
                    classes=classes)
    out = dict(
        input_joints=joints.astype(np.float32),
        box_label_mask=np.zeros(MAX_GT, np.float32),
        sem_cls_label=np.zeros(MAX_GT, np.int64),
        center_label=np.zeros((MAX_GT, 3), np.float32),
        size=np.zeros((MAX_GT, 3), np.float32),
        heading=np.zeros((MAX_GT, 2), np.float32),
        vote_label=votes.astype(np.float32),
        vote_label_mask=mask,
    )
    out["box_label_mask"][:n_gt] = 1
    out["sem_cls_label"][:n_gt] = classes
    out["center_label"][:n_gt] = centers
    out["size"][:n_gt] = np.log(sizes)
    out["heading"][:n_gt, 0] = np.sin(theta)
    out["heading"][:n_gt, 1] = np.cos(theta)
    return out


def make_raw_sample(rng, F, J, name="synthetic"):
    """One sample in the reference's ON-DISK schema (what utils/virtualhome/3_generate_samples.py:188-193
    writes and dataloader.py:90-101 reads back): skeleton_joints f32 (F,J,3), skeleton_joint_votes f32
    (F,J,10) with column 0 = vote mask, object_nodes = list of dict(class_id, centroid f32 (3,), R_mat f32
    (3,3) with rows (heading, up, right), size f32 (3,))."""
    sc = make_scene(rng, F, J, raw=True)
    votes10 = np.concatenate([sc["mask"][..., None].astype(np.float64), sc["votes"]], -1)
    nodes = [dict(class_id=np.int32(sc["classes"][g]), centroid=sc["centers"][g].astype(np.float32),
                  R_mat=_rot(sc["theta"][g]).astype(np.float32), size=sc["sizes"][g].astype(np.float32))
             for g in range(len(sc["theta"]))]
    return dict(skeleton_joints=sc["joints"].astype(np.float32),
                skeleton_joint_votes=votes10.astype(np.float32), object_nodes=nodes, name=name)


def make_batch(B, T, J, seed=1234, as_torch=True, pin=False):
    """A collated batch of B scenes.  seed = 1234 + rank by convention."""
    rng = np.random.default_rng(seed)
    scenes = [make_scene(rng, T, J) for _ in range(B)]
    batch = {k: np.stack([s[k] for s in scenes]) for k in scenes[0]}
    batch["sample_idx"] = ["synthetic_%d_%d" % (seed, i) for i in range(B)]
    if as_torch:
        import torch
        for k, v in list(batch.items()):
            if isinstance(v, np.ndarray):
                t = torch.from_numpy(v)
                batch[k] = t.pin_memory() if pin else t
    return batch


def make_cloud(B, N, seed=0, extent=2.0):
    """Microbench point clouds (B,N,3) f32: a random walk blurred by N(0,0.3^2), like T*J joints."""
    rng = np.random.default_rng(seed)
    walk = np.cumsum(rng.normal(0.0, 0.05, size=(B, N, 3)), axis=1)
    pts = walk + rng.normal(0.0, 0.3, size=(B, N, 3))
    return np.clip(pts, -extent * 4, extent * 4).astype(np.float32)


def deterministic_state_dict(reference_state_dict, seed=0):
    """Platform-independent synthetic weights for a P2RNet state-dict (there are no checkpoints offline).

    Every tensor is re-drawn from a torch CPU generator seeded by (seed, position of the key in sorted
    order), with fan-in scaling for conv weights, so any process that knows the key -> shape/dtype map (the
    reference model in the build container, the B200 model on the GPU box, the CPU checker) builds the SAME
    weights without shipping megabytes of fixtures.  The GMM mixing logits get bias -4.6 and 0.1x weights so
    that sum(pi) ~ 1 and decoded boxes are plausible ("Eval with untrained weights aborts").
    mu tensors are kept (they come from the reference's grid initialisation and are stored with the goldens).
    """
    import torch
    out = {}
    for i, k in enumerate(sorted(reference_state_dict)):
        v = reference_state_dict[k]
        g = torch.Generator().manual_seed(1_000_003 * (seed + 1) + i)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k == "backbone.A" or k.endswith(".mdn.mu"):
            out[k] = v.clone()
        elif k.endswith(".mdn.log_sigma"):
            out[k] = torch.full_like(v, -1.0)
        elif k.endswith("running_mean"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith("running_var"):
            out[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif "edge_importance" in k:
            out[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif "batchnorm" in k or ".tcn.0." in k or ".tcn.3." in k:
            out[k] = (1.0 if k.endswith("weight") else 0.0) + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith(".mdn.pi.conv.weight"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g) / float(v[0].numel()) ** 0.5
        elif k.endswith(".mdn.pi.conv.bias"):
            out[k] = -4.6 + 0.05 * torch.randn(v.shape, generator=g)
        elif k.endswith("conv_sem_obj.2.conv.weight"):
            # wide objectness / class logits: NMS order must not hinge on 1-ulp score differences
            out[k] = 8.0 * torch.randn(v.shape, generator=g) / float(v[0].numel()) ** 0.5
        elif k.endswith("bias"):
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        else:  # conv / linear weights
            out[k] = torch.randn(v.shape, generator=g) / float(v[0].numel()) ** 0.5
        out[k] = out[k].to(v.dtype)
    return out


# ---------------------------------------------------------------------------------------------------------------
# Evaluation set of BASELINE.json config #5 / SURVEY.md section 8(d): scenes whose K = 128 predictions are drawn as
# ground truth + noise, so that NMS has duplicates to remove, the far-box filter has something to drop and mAP is far
# from both 0 and 1.  Every scene is generated from (seed, scene index) alone: any subset can be rebuilt anywhere.
def make_eval_scene(seed, index, T=256, K=128):
    """-> dict of numpy arrays: input_joints f32 (T,1,3) (the hip joint is all parse_predictions reads), the ground-truth
    label arrays of make_scene, and the network-output arrays center f32 (K,3), size f32 (K,3) = log size, heading f64
    (K,2) = (sin, cos) (the reference's heading head is float64), objectness_scores f32 (K,2), sem_cls_scores f32 (K,22)."""
    rng = np.random.default_rng([int(seed), int(index)])
    hip = np.cumsum(rng.normal(0.0, 0.05, size=(T, 3)), axis=0)
    hip[:, 1] = np.clip(0.9 + hip[:, 1], 0.8, 1.0)
    n_gt = int(rng.integers(1, MAX_GT + 1))
    g_center = hip[rng.integers(0, T, size=n_gt)] + rng.normal(0.0, 0.3, size=(n_gt, 3))
    g_size = rng.uniform(0.2, 1.7, size=(n_gt, 3))
    g_theta = rng.uniform(-np.pi, np.pi, size=n_gt)
    g_cls = rng.integers(0, NUM_CLASS, size=n_gt)

    which = rng.integers(0, n_gt, size=K)
    sigma = rng.choice([0.03, 0.1, 0.25, 0.5], size=K)
    junk = rng.random(K) < 0.15
    center = g_center[which] + sigma[:, None] * rng.normal(size=(K, 3))
    log_size = np.log(g_size[which]) + 0.6 * sigma[:, None] * rng.normal(size=(K, 3))
    theta = g_theta[which] + 0.8 * sigma * rng.normal(size=K)
    center[junk] = hip[rng.integers(0, T, size=int(junk.sum()))] + rng.normal(0.0, 1.0, size=(int(junk.sum()), 3))
    log_size[junk] = np.log(rng.uniform(0.1, 2.5, size=(int(junk.sum()), 3)))
    theta[junk] = rng.uniform(-np.pi, np.pi, size=int(junk.sum()))
    center[0:3] += 30.0                         # far from the trajectory: removed by remove_far_box
    log_size[3:5, 0] = np.log(0.004)            # degenerate / absurd sizes: removed
    log_size[5, 1] = np.log(14.0)
    radius = 1.0 + 0.1 * rng.normal(size=K)     # (sin, cos) off the unit circle, like a regressed heading
    objectness = np.zeros((K, 2))
    objectness[:, 1] = np.where(junk, -1.0, 3.0 - 8.0 * sigma) + 0.7 * rng.normal(size=K)
    sem = rng.normal(size=(K, NUM_CLASS))
    right = rng.random(K) < 0.8
    boosted = np.where(right & ~junk, g_cls[which], rng.integers(0, NUM_CLASS, size=K))
    sem[np.arange(K), boosted] += 4.0

    out = dict(
        input_joints=hip[:, None, :].astype(np.float32),
        box_label_mask=np.zeros(MAX_GT, np.float32), sem_cls_label=np.zeros(MAX_GT, np.int64),
        center_label=np.zeros((MAX_GT, 3), np.float32), size=np.zeros((MAX_GT, 3), np.float32),
        heading=np.zeros((MAX_GT, 2), np.float32),
        pred_center=center.astype(np.float32), pred_size=log_size.astype(np.float32),
        pred_heading=np.stack([radius * np.sin(theta), radius * np.cos(theta)], -1),
        pred_objectness_scores=objectness.astype(np.float32), pred_sem_cls_scores=sem.astype(np.float32))
    out["box_label_mask"][:n_gt] = 1
    out["sem_cls_label"][:n_gt] = g_cls
    out["center_label"][:n_gt] = g_center
    out["size"][:n_gt] = np.log(g_size)
    out["heading"][:n_gt, 0] = np.sin(g_theta)
    out["heading"][:n_gt, 1] = np.cos(g_theta)
    return out


def make_eval_batch(seed, start, count, T=256, K=128):
    """Scenes [start, start + count) collated -> (est_data, gt_data) dicts of torch CPU tensors in the schema
    parse_predictions / parse_groundtruths read (net_utils/ap_helper.py:133-292)."""
    import torch
    scenes = [make_eval_scene(seed, i, T, K) for i in range(start, start + count)]
    stack = lambda k: torch.from_numpy(np.stack([s[k] for s in scenes]))
    est = {k: stack("pred_" + k) for k in ("center", "size", "heading", "objectness_scores", "sem_cls_scores")}
    gt = {k: stack(k) for k in ("input_joints", "box_label_mask", "sem_cls_label", "center_label", "size", "heading")}
    return est, gt
