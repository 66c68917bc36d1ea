"""Build the UNMODIFIED reference pointnet2 `_ext` for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  The sources are compiled where
they lie under /root/reference (external/pointnet2_ops_lib/pointnet2_ops/_ext-src);
nothing is copied into the repo, only the resulting .so lands in oracle/_ref/
(git-ignored, but it travels to the GPU box with the gpurun snapshot).  On the GPU
box the .so is the index-exactness referee for FPS / ball_query / three_nn and the
"reference kernels on B200" timing baseline.

The reference's own arch list (setup.py:19, pointnet2_utils.py:23) stops at sm_75
and contains 3.7, which nvcc 12.9 rejects, so the arch list is overridden here.
"""
import glob
import os
import os.path as osp
import sys

REF_SRC = "/root/reference/external/pointnet2_ops_lib/pointnet2_ops/_ext-src"
OUT_DIR = osp.join(osp.dirname(osp.abspath(__file__)), "_ref")
NAME = "p2r_ref_ext"


def so_path():
    return osp.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    if not osp.isdir(REF_SRC):
        return None
    if osp.exists(so_path()):
        return so_path()
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    from torch.utils.cpp_extension import load
    srcs = sorted(glob.glob(osp.join(REF_SRC, "src", "*.cpp")) +
                  glob.glob(osp.join(REF_SRC, "src", "*.cu")))
    load(NAME, sources=srcs, extra_include_paths=[osp.join(REF_SRC, "include")],
         extra_cflags=["-O3"], extra_cuda_cflags=["-O3", "-lineinfo"],
         build_directory=OUT_DIR, with_cuda=True, verbose=verbose, is_python_module=False)
    return so_path()


def load_ref_ext():
    """Import the prebuilt reference extension (GPU box or here). Returns None if absent."""
    p = so_path()
    if not osp.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
