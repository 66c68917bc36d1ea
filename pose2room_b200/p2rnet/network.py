"""P2RNet: backbone -> centre voting -> proposal / box heads, with the reference's model API.

Mirrors /root/reference/models/p2rnet/modules/network.py:10-106 and the parts of BaseNetwork it relies on
(models/network.py:9-81): `net(data) -> end_points`, `net.generate(data, eval=True) -> (end_points, eval_dict,
parsed_predictions)`, `net.loss(pred_or_tuple, data) -> dict`, `set_mode()`, `load_weight()` (strips the
leading 'module.'), sub-modules built from cfg.config['model'] through the MODULES / LOSSES registries and
exposed as named children backbone / centervoting / detection with `<phase>_loss` attributes.
"""
import torch
import torch.nn as nn

from .. import ap_helper
from .registers import LOSSES, METHODS, MODULES
from .vote_center import fused_vote_enabled


@METHODS.register_module
class P2RNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        phase_names = []
        if cfg.config[cfg.config["mode"]]["phase"] in ["full"]:
            phase_names += ["backbone", "centervoting", "detection"]
        if (not cfg.config["model"]) or (not phase_names):
            cfg.log_string("No submodule found. Please check the phase name and model definition.")
            raise ModuleNotFoundError("No submodule found. Please check the phase name and model definition.")
        for phase_name, net_spec in cfg.config["model"].items():
            if phase_name not in phase_names:
                continue
            optim_spec = self.load_optim_spec(cfg.config, net_spec)
            self.add_module(phase_name, MODULES.get(net_spec["method"])(cfg, optim_spec))
            loss_cls = LOSSES.get(net_spec["loss"], "Null")
            setattr(self, phase_name + "_loss", loss_cls(net_spec.get("weight", 1), cfg.config["device"]["gpu"], cfg))
        self.freeze_modules(cfg)

    # ---- BaseNetwork behaviour ----------------------------------------------------------------
    def load_optim_spec(self, config, net_spec):
        if config["mode"] != "train":
            return None
        if "optimizer" in net_spec:
            spec = config["optimizer"].copy()
            for key in spec:
                spec[key] = net_spec["optimizer"].get(key, spec[key])
            return spec
        return config["optimizer"]

    def freeze_modules(self, cfg):
        if cfg.config["mode"] != "train":
            return
        for layer in cfg.config["train"]["freeze"]:
            mod = self
            try:
                for part in layer.split("."):
                    mod = getattr(mod, part)
            except AttributeError:
                continue
            for p in mod.parameters():
                p.requires_grad = False
            cfg.log_string("The module: %s is fixed." % layer)

    def set_mode(self):
        freeze = self.cfg.config["train"]["freeze"]
        for name, child in self.named_children():
            if name in freeze:
                child.train(False)

    def load_weight(self, pretrained_model):
        model_dict = self.state_dict()
        stripped = {".".join(k.split(".")[1:]): v for k, v in pretrained_model.items()}
        picked = {k: v for k, v in stripped.items() if k in model_dict}
        missing = set(k.split(".")[0] for k in model_dict if k not in picked)
        self.cfg.log_string(str(missing) + " subnet missed.")
        model_dict.update(picked)
        self.load_state_dict(model_dict)

    # ---- model API ------------------------------------------------------------------------------
    def _trunk(self, data):
        end_points = self.backbone(data["input_joints"], {})
        if fused_vote_enabled():     # residual adds + the normalisation below as one kernel (opt-in, vote_center.py)
            xyz, features = self.centervoting(end_points["seed_skeleton"], end_points["seed_features"], normalize=True)
        else:
            xyz, features = self.centervoting(end_points["seed_skeleton"], end_points["seed_features"])
            features = features.div(torch.norm(features, p=2, dim=2).unsqueeze(2))   # network.py:89-90
        end_points["vote_xyz"] = xyz
        end_points["vote_features"] = features
        return end_points, xyz, features

    def forward(self, data):
        end_points, xyz, features = self._trunk(data)
        end_points, _ = self.detection(xyz, features, end_points, False)
        return end_points

    def generate(self, data, eval=True):
        end_points, xyz, features = self._trunk(data)
        end_points, _ = self.detection.generate(xyz, features, end_points, False)
        eval_dict, parsed = ap_helper.parse_predictions(end_points, data, self.cfg.eval_config)
        eval_dict = ap_helper.assembly_pred_map_cls(eval_dict, parsed, self.cfg.eval_config)
        if eval:
            parsed_gts = ap_helper.parse_groundtruths(data, self.cfg.eval_config)
            eval_dict["batch_gt_map_cls"] = ap_helper.assembly_gt_map_cls(parsed_gts)
        return end_points, eval_dict, parsed

    def loss(self, pred_data, gt_data):
        if isinstance(pred_data, tuple):
            pred_data = pred_data[0]
        return self.detection_loss(pred_data, gt_data, self.cfg.dataset_config)
