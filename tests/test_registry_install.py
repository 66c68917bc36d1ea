"""The two bindings INTEGRATION.md tells a maintainer to call, exercised:

  * registers.install_into(METHODS, MODULES, LOSSES): the reference's OWN registries then build this package's P2RNet from
    the reference's OWN YAML + mount_external_config, through the reference's own factory calls -- `load_model`'s lookup
    (net_utils/utils.py:247), `load_optimizer` (models/optimizers.py:60-100: optim_spec plumbing -> AdamW) and
    `load_trainer` (net_utils/utils.py:257-270 -> models/p2rnet/config.py -> Trainer) -- in the build container, CPU
    (`needs_reference`: construction / registry / optimiser / trainer wiring; the forward needs a GPU);
  * ext.install_as_pointnet2_ops(): the UNMODIFIED reference `pointnet2_utils` / `pointnet2_modules` import on top of this
    package's native operator module (fresh interpreter, because the binding must precede the first import);
  * on the GPU box: the reference's `BaseTrainer.train_step` / `Trainer.compute_loss` call sequence
    (models/training.py:25-43, models/p2rnet/training.py:100-121), replicated line for line around `nn.DataParallel(net)`
    as `load_model` wraps it (utils.py:253), drives the product for two optimiser steps.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.needs_reference
def test_install_into_reference_registries_and_build_through_the_reference_factories():
    from oracle import ref_import
    from pose2room_b200 import p2rnet
    from pose2room_b200.p2rnet import registers
    ns = ref_import.import_reference()
    saved = [dict(r.module_dict) for r in (ns.METHODS, ns.MODULES, ns.LOSSES)]
    cwd = os.getcwd()
    try:
        registers.install_into(ns.METHODS, ns.MODULES, ns.LOSSES)
        for name in ("STGCN", "CenterVoteModule", "ProposalNet"):
            assert ns.MODULES.get(name) is registers.MODULES.get(name)
        assert ns.LOSSES.get("BoxNetDetectionLoss") is registers.LOSSES.get("BoxNetDetectionLoss")
        assert ns.LOSSES.get("no such loss", "Null") is registers.LOSSES.get("Null")
        os.chdir(ns.scratch)
        config = ns.config_utils.read_to_dict(os.path.join(ref_import.REF_ROOT, "configs/config_files/p2rnet_train.yaml"))
        config["mode"] = "train"
        config["device"].update(distributed=False, is_main_process=True, gpu=0)
        cfg = ref_import._Cfg(config)
        ns.config_utils.mount_external_config(cfg)              # the reference's Dataset_Config (53 joints)
        net = ns.METHODS.get(cfg.config["method"])(cfg)          # utils.py:247
        assert type(net) is p2rnet.P2RNet
        assert [n for n, _ in net.named_children()] == ["backbone", "centervoting", "detection"]     # train.py:56-57
        assert isinstance(net.detection_loss, registers.LOSSES.get("BoxNetDetectionLoss"))
        assert net.backbone.A.shape == (11, 53, 53) and len(net.state_dict()) == 219
        import models.optimizers as ref_opt
        optimizer = ref_opt.load_optimizer(cfg.config, net)     # optim_spec on every phase module (optimizers.py:22-39)
        assert isinstance(optimizer, torch.optim.AdamW)
        n_opt = sum(p.numel() for g in optimizer.param_groups for p in g["params"])
        assert n_opt == sum(p.numel() for p in net.parameters() if p.requires_grad)
        scheduler = ref_opt.load_scheduler(cfg.config, optimizer)
        assert scheduler.milestones
        import net_utils.utils as ref_utils
        wrapped = torch.nn.DataParallel(net, device_ids=[0]) if torch.cuda.is_available() else _Wrapper(net)
        trainer = ref_utils.load_trainer(cfg, wrapped, optimizer, torch.device("cpu"))
        assert type(trainer).__module__ == "models.p2rnet.training" and trainer.net.module is net
        assert callable(trainer.net.module.loss) and callable(trainer.net.module.generate)
        # the reference's checkpoint loader on a reference-made state dict (network.py:59-67: strips 'module.')
        ref_net, _ = ref_import.build_reference_model(mode="train", joint_num=53, num_frames=768)
        net.load_weight({"module." + k: v for k, v in ref_net.state_dict().items()})
        for k, v in ref_net.state_dict().items():
            assert torch.equal(net.state_dict()[k], v), k
    finally:
        os.chdir(cwd)
        for r, s in zip((ns.METHODS, ns.MODULES, ns.LOSSES), saved):
            r.module_dict.clear()
            r.module_dict.update(s)


class _Wrapper(torch.nn.Module):
    """`.module` holder for the CPU-only container (nn.DataParallel refuses to build without a GPU)."""

    def __init__(self, module):
        super().__init__()
        self.module = module


@pytest.mark.needs_reference
def test_install_as_pointnet2_ops_feeds_the_unmodified_reference_python():
    code = r'''
import sys, types
for name in ["h5py", "trimesh", "matplotlib", "seaborn", "vtk", "plyfile"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference")       # the reference runs from its repository root (main.py)
sys.path.insert(0, %r)
from pose2room_b200 import ext
REF_PKG = "/root/reference/external/pointnet2_ops_lib/pointnet2_ops"
ext.install_as_pointnet2_ops(REF_PKG)
# the reference's files, unmodified, imported the way the reference imports them (proposal_net.py:11, pointnet2_modules.py:7)
from external.pointnet2_ops_lib.pointnet2_ops import pointnet2_utils, pointnet2_modules
assert pointnet2_utils.__file__.startswith("/root/reference/"), pointnet2_utils.__file__
assert pointnet2_utils._ext is ext
for name in ext.NAMES:
    assert callable(getattr(pointnet2_utils._ext, name)), name
sa = pointnet2_modules.PointnetSAModuleVotes(npoint=128, radius=0.3, nsample=16, mlp=[256, 256, 256, 256], use_xyz=False,
                                             normalize_xyz=True, bn=False)
assert sa.grouper.radius == 0.3
import torch
try:
    pointnet2_utils.furthest_point_sample(torch.zeros(1, 8, 3), 4)
    raise SystemExit("a CPU tensor must be refused")
except RuntimeError as e:
    assert "CPU not supported" in str(e), e
print("INSTALL-OK")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "INSTALL-OK" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])


class _ReplicaTrainer:
    """models/training.py:12-43 (BaseTrainer) + models/p2rnet/training.py:100-121 (Trainer.to_device / compute_loss),
    the statements in the reference's order; reduce_dict (utils.py:490-514) is the identity on one process."""

    def __init__(self, cfg, net, optimizer, device=None):
        self.cfg, self.net, self.optimizer, self.device = cfg, net, optimizer, device

    def to_device(self, data):
        for key in data:
            if key in ["sample_idx"]:
                continue
            data[key] = data[key].to(self.device)
        return data

    def compute_loss(self, data):
        data = self.to_device(data)
        est_data = self.net(data)
        return self.net.module.loss(est_data, data)

    def train_step(self, data):
        self.optimizer.zero_grad()
        loss = self.compute_loss(data)
        if loss["total"].requires_grad:
            loss["total"].backward()
            if self.cfg.config["optimizer"]["clip_norm"] > 0:
                torch.nn.utils.clip_grad_norm_(self.net.parameters(), self.cfg.config["optimizer"]["clip_norm"])
            self.optimizer.step()
        return {k: v.item() for k, v in loss.items()}


@pytest.mark.gpu
def test_reference_train_step_sequence_drives_the_product(cuda):
    from pose2room_b200 import synthetic
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = P2RConfig(mode="train", joint_num=25, num_frames=160, num_seeds=64, num_target=16)
    net = P2RNet(cfg)
    net.load_state_dict(synthetic.deterministic_state_dict(net.state_dict(), seed=7))
    net = torch.nn.DataParallel(net.to(cuda), device_ids=[0])                # load_model, utils.py:250-253
    optimizer = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3)
    trainer = _ReplicaTrainer(cfg, net, optimizer, cuda)
    net.train()
    net.module.set_mode()                                                    # train_epoch.py:31-32
    before = {k: v.detach().clone() for k, v in net.module.named_parameters()}
    torch.manual_seed(3)
    first = trainer.train_step(synthetic.make_batch(2, 160, 25, seed=11))   # host tensors: to_device moves them
    second = trainer.train_step(synthetic.make_batch(2, 160, 25, seed=11))
    assert set(first) == {"total", "vote_loss", "objectness_loss", "center_loss", "size_loss", "heading_loss",
                          "sem_cls_loss", "pos_ratio", "neg_ratio", "obj_acc"}
    assert all(np.isfinite(v) for v in first.values()) and all(np.isfinite(v) for v in second.values())
    moved = sum(int(not torch.equal(before[k], v)) for k, v in net.module.named_parameters())
    assert moved >= 0.9 * len(before), (moved, len(before))
    # the same first step without the trainer / wrapper: same loss
    torch.manual_seed(0)
    np.random.seed(0)
    net2 = P2RNet(cfg)
    net2.load_state_dict(synthetic.deterministic_state_dict(net2.state_dict(), seed=7))
    net2 = net2.to(cuda).train()
    data = {k: (v.to(cuda) if isinstance(v, torch.Tensor) else v) for k, v in synthetic.make_batch(2, 160, 25, seed=11).items()}
    torch.manual_seed(3)
    direct = net2.loss(net2(data), data)["total"].item()
    assert abs(direct - first["total"]) <= 1e-5 * abs(direct), (direct, first["total"])
