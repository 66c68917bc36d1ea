"""Model-level goldens: run the UNMODIFIED reference P2RNet (forward / loss / backward / generate) on CPU,
with the C oracle plugged in as pointnet2_ops._ext, and store compact outputs.

Weights are not stored: both sides rebuild them with pose2room_b200.synthetic.deterministic_state_dict
(only the GMM `mu` grids, which the reference initialises with a randomly-started FPS, are stored).
Called from tests/golden/make_golden.py (`python tests/golden/make_golden.py model`).
"""
import os.path as osp

import numpy as np
import torch

from oracle import pointnet2_ref, ref_import
from pose2room_b200 import synthetic

OUT = osp.dirname(osp.abspath(__file__))

CONFIGS = {
    # name: (B, T, J, num_seeds, num_target)
    "small": (2, 160, 25, 64, 16),
    "ref53": (1, 96, 53, 128, 32),      # T < num_seeds -> the linspace seed branch, reference 53-joint rig
    "bl": (1, 1024, 25, 512, 128),      # one sequence at the BASELINE shape
}
GRAD_KEYS = ["backbone.sk_feat.0.conv.weight", "backbone.pos_embed.2.conv.bias", "backbone.edge_importance.3",
             "backbone.st_gcn_networks.0.gcn.conv.bias", "backbone.st_gcn_networks.5.tcn.2.weight",
             "backbone.st_gcn_networks.2.tcn.3.weight", "centervoting.conv_input.2.conv.bias",
             "detection.vote_aggregation.mlp_module.0.bias", "detection.conv_sem_obj.2.conv.weight",
             "detection.gmm_size.mdn.pi.conv.bias", "detection.gmm_heading.mdn.mu"]


def build(mode, J, T, S, P):
    net, cfg = ref_import.build_reference_model(mode=mode, joint_num=J, num_frames=T, ext=pointnet2_ref.RefExt)
    return net, cfg


def run_config(name, out):
    B, T, J, S, P = CONFIGS[name]
    data = synthetic.make_batch(B, T, J, seed=1234)
    # --- training-mode forward + loss + backward -------------------------------------------------
    import configs.config_utils  # noqa  (reference already importable)
    net, cfg = ref_import.build_reference_model(mode="train", joint_num=J, num_frames=T, ext=pointnet2_ref.RefExt)
    cfg.config["data"]["num_seeds"], cfg.config["data"]["num_target"] = S, P
    net, cfg = _rebuild(cfg, J, "train")
    sd = synthetic.deterministic_state_dict(net.state_dict(), seed=7)
    for k in sd:
        if k.endswith(".mdn.log_sigma"):
            sd[k] = torch.full_like(sd[k], -50.0)     # sigma ~ 2e-22: the train-mode noise term vanishes
    net.load_state_dict(sd)
    net.train()
    for k in sd:
        if k.endswith(".mdn.mu"):
            out["%s_%s" % (name, k)] = sd[k].numpy()
    ep = net(data)
    loss = net.loss(ep, data)
    loss["total"].backward()
    for k in ["seed_inds", "aggregated_vote_inds", "vote_xyz", "aggregated_vote_xyz", "center", "size", "heading",
              "objectness_scores", "sem_cls_scores"]:
        out["%s_train_%s" % (name, k)] = ep[k].detach().numpy()
    for k in ["seed_features", "vote_features"]:
        t = ep[k].detach().double()
        out["%s_train_%s_stats" % (name, k)] = np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])
        out["%s_train_%s_head" % (name, k)] = ep[k].detach()[:, :4, :32].numpy()
    for k, v in loss.items():
        out["%s_loss_%s" % (name, k)] = np.array(v.item())
    params = dict(net.named_parameters())
    for k in GRAD_KEYS:
        out["%s_grad_%s" % (name, k)] = params[k].grad.numpy()
    gn = {k: (p.grad.double().norm().item() if p.grad is not None else -1.0) for k, p in params.items()}
    out["%s_gradnorm_keys" % name] = np.array(sorted(gn))
    out["%s_gradnorm_vals" % name] = np.array([gn[k] for k in sorted(gn)])
    # running statistics after one training step (BatchNorm side effect)
    st = net.state_dict()
    for k in ["backbone.st_gcn_networks.3.tcn.0.running_mean", "backbone.st_gcn_networks.3.tcn.3.running_var",
              "detection.conv_size.1.batchnorm.running_mean"]:
        out["%s_after_%s" % (name, k)] = st[k].numpy()
    # --- eval-mode generate (deterministic) ------------------------------------------------------
    net_t, cfg_t = _rebuild(_cfg_for("test", J, T, S, P), J, "test")
    net_t.load_state_dict(sd)
    net_t.eval()
    with torch.no_grad():
        ep, eval_dict, parsed = net_t.generate(data)
    for k in ["seed_inds", "aggregated_vote_inds", "vote_xyz", "aggregated_vote_xyz", "center", "size", "heading",
              "objectness_scores", "sem_cls_scores"]:
        out["%s_gen_%s" % (name, k)] = ep[k].numpy()
    out["%s_gen_pred_mask" % name] = eval_dict["pred_mask"]
    out["%s_gen_corners" % name] = parsed["pred_corners_3d"]
    out["%s_gen_obj_prob" % name] = parsed["obj_prob"]
    out["%s_gen_npred" % name] = np.array([len(x) for x in eval_dict["batch_pred_map_cls"]])
    print(name, "loss", float(loss["total"]), "picked", eval_dict["pred_mask"].sum(axis=1))


def _cfg_for(mode, J, T, S, P):
    _, cfg = ref_import.build_reference_model(mode=mode, joint_num=J, num_frames=T, ext=pointnet2_ref.RefExt)
    cfg.config["data"]["num_seeds"], cfg.config["data"]["num_target"] = S, P
    return cfg


def _rebuild(cfg, J, mode):
    """Re-instantiate the reference P2RNet from an edited cfg (num_seeds / num_target overrides)."""
    import os
    ns = ref_import.import_reference()
    import models.p2rnet.modules.stgcn as ref_stgcn
    orig = ref_stgcn.Graph
    if J == 25:
        ref_stgcn.Graph = lambda layout="virtualroom", **kw: orig(layout="ntu-rgb+d", **kw)
    cwd = os.getcwd()
    os.chdir(ns.scratch)
    try:
        torch.manual_seed(42)
        np.random.seed(42)
        net = ns.METHODS.get("P2RNet")(cfg)
    finally:
        ref_stgcn.Graph = orig
        os.chdir(cwd)
    return net, cfg


def main(ns=None):
    out = {}
    for name in CONFIGS:
        run_config(name, out)
    np.savez_compressed(osp.join(OUT, "p2rnet.npz"), **out)
    print("p2rnet.npz:", len(out), "arrays")
