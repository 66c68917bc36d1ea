"""Import the UNMODIFIED reference (yinyunie/Pose2Room) from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY.  Used in THIS container (where /root/reference exists) to
  * validate the oracle restatements in this directory against the real reference, and
  * generate the golden vectors committed under tests/golden/ (tests/golden/make_golden.py).
Nothing under `pose2room_b200/` imports this file, and nothing run on the GPU box needs
/root/reference (it does not exist there).

Recipe = SURVEY.md Appendix C:
  1. chdir to a scratch dir (Dataset_Config mkdirs relative paths at import,
     configs/dataset_config.py:67-76; the reference tree is read-only);
  2. stub the six absent third-party modules (h5py, trimesh, matplotlib, seaborn, plyfile, vtk);
  3. provide `pointnet2_ops` with an `_ext` attribute BEFORE importing `models`
     (pointnet2_utils.py:7-8 imports `pointnet2_ops._ext`);
  4. import net_utils.utils before models (circular import, main.py:22 order).
"""
import os
import sys
import tempfile
import types

REF_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_DONE = {}


def import_reference(ext=None):
    """Make the reference importable; `ext` is the object exposed as pointnet2_ops._ext
    (nine functions).  Returns a namespace with the interesting reference modules."""
    if "ns" in _DONE:
        if ext is not None:
            set_ext(ext)
        return _DONE["ns"]
    assert available(), "reference tree not present"
    scratch = tempfile.mkdtemp(prefix="p2r_refscratch_")
    os.makedirs(os.path.join(scratch, "datasets"), exist_ok=True)
    _DONE["cwd"] = os.getcwd()
    os.chdir(scratch)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    for name in ["h5py", "trimesh", "trimesh.exchange", "matplotlib", "seaborn", "vtk"]:
        if name not in sys.modules:
            _stub(name)
    _stub("trimesh.exchange.binvox", voxelize_mesh=None)
    _stub("matplotlib.pyplot")
    _stub("matplotlib.lines", Line2D=object)
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("vtk.util") if "vtk.util" not in sys.modules else None
    _stub("vtk.util.numpy_support", numpy_to_vtk=None, vtk_to_numpy=None)

    pkg = types.ModuleType("pointnet2_ops")
    pkg.__path__ = [os.path.join(REF_ROOT, "external/pointnet2_ops_lib/pointnet2_ops")]
    sys.modules["pointnet2_ops"] = pkg
    holder = types.ModuleType("pointnet2_ops._ext")
    sys.modules["pointnet2_ops._ext"] = holder
    pkg._ext = holder
    _DONE["ext_holder"] = holder
    if ext is not None:
        set_ext(ext)

    import net_utils.utils  # noqa: F401  (must precede `models`)
    import models  # noqa: F401
    import models.loss
    import models.p2rnet.modules.network
    import net_utils.nn_distance
    import net_utils.nms
    import net_utils.box_util
    import net_utils.ap_helper
    import net_utils.eval_det
    import net_utils.vn_dgcnn_util
    import net_utils.libs
    import utils.pc_utils
    import utils.tools
    import configs.config_utils
    from external.pointnet2_ops_lib.pointnet2_ops import pointnet2_utils, pointnet2_modules
    from models.registers import METHODS, MODULES, LOSSES

    ns = types.SimpleNamespace(
        nn_distance=net_utils.nn_distance, nms=net_utils.nms, box_util=net_utils.box_util,
        ap_helper=net_utils.ap_helper, eval_det=net_utils.eval_det,
        vn_dgcnn_util=net_utils.vn_dgcnn_util, libs=net_utils.libs, pc_utils=utils.pc_utils,
        tools=utils.tools, config_utils=configs.config_utils, loss=models.loss,
        pointnet2_utils=pointnet2_utils, pointnet2_modules=pointnet2_modules,
        METHODS=METHODS, MODULES=MODULES, LOSSES=LOSSES, scratch=scratch)
    _DONE["ns"] = ns
    os.chdir(_DONE["cwd"])
    return ns


def set_ext(ext):
    """(Re)bind the nine native entry points seen by the reference's pointnet2_utils."""
    holder = _DONE["ext_holder"]
    for name in ["furthest_point_sampling", "gather_points", "gather_points_grad", "three_nn",
                 "three_interpolate", "three_interpolate_grad", "ball_query", "group_points",
                 "group_points_grad"]:
        setattr(holder, name, getattr(ext, name))


class _Cfg:
    """Stand-in for configs.config_utils.CONFIG (which would create log dirs,
    config_utils.py:78-97): anything with .config and .log_string works."""

    def __init__(self, config):
        self.config = config

    def log_string(self, *a, **k):
        pass


def build_reference_model(mode="train", joint_num=25, num_frames=1024, seed=42, ext=None,
                          plausible_gmm=False):
    """Reference P2RNet on CPU with J/T overridden."""
    import torch
    ns = import_reference(ext)
    cwd = os.getcwd()
    os.chdir(ns.scratch)
    try:
        yaml_name = "p2rnet_train.yaml" if mode == "train" else "p2rnet_test.yaml"
        config = ns.config_utils.read_to_dict(os.path.join(REF_ROOT, "configs/config_files", yaml_name))
        config["mode"] = mode
        config["device"].update(distributed=False, is_main_process=True, gpu="cpu")
        config["data"]["num_frames"] = num_frames
        if mode != "train":
            config["train"] = {"freeze": []}
        cfg = _Cfg(config)
        ns.config_utils.mount_external_config(cfg)
        cfg.dataset_config.joint_num = joint_num
        import models.p2rnet.modules.stgcn as ref_stgcn
        orig_graph = ref_stgcn.Graph
        if joint_num == 25:
            def graph25(layout="virtualroom", **kw):
                return orig_graph(layout="ntu-rgb+d", **kw)
            ref_stgcn.Graph = graph25
        elif joint_num != 53:
            raise ValueError("reference has 25- and 53-joint graph layouts only")
        try:
            torch.manual_seed(seed)
            import numpy as np
            np.random.seed(seed)
            net = ns.METHODS.get("P2RNet")(cfg)
        finally:
            ref_stgcn.Graph = orig_graph
        if plausible_gmm:
            # SURVEY.md §7 hard parts: default init empties every scene in remove_far_box.
            with torch.no_grad():
                for g in [net.detection.gmm_center, net.detection.gmm_size, net.detection.gmm_heading]:
                    g.mdn.pi.conv.bias.fill_(-4.6)
                    g.mdn.pi.conv.weight.mul_(0.1)
    finally:
        os.chdir(cwd)
    return net, cfg
