"""Build libp2r_b200.so in-tree with nvcc for sm_100a (no torch, no pybind: a plain C-ABI library).

    python -m pose2room_b200.build [--force] [-v]

The .so lands in pose2room_b200/lib/ (git-ignored, but it travels to the GPU box with gpurun).
"""
import hashlib
import os
import os.path as osp
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = osp.dirname(osp.abspath(__file__))
CSRC = osp.join(HERE, "csrc")
LIBDIR = osp.join(HERE, "lib")
OBJDIR = osp.join(HERE, "build")
LIB = osp.join(LIBDIR, "libp2r_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-I", osp.join(osp.dirname(HERE), "include"), "-I", CSRC]
# per-file extras: geometry_ops keeps fp64 polygon arithmetic un-contracted (numpy has no FMA)
EXTRA = {"geometry_ops.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")) or osp.join(CSRC, f) == path:
            h.update(open(osp.join(CSRC, f), "rb").read())
    h.update(open(osp.join(osp.dirname(HERE), "include", "p2r_b200.h"), "rb").read())
    h.update(" ".join(COMMON + EXTRA.get(osp.basename(path), [])).encode())
    return h.hexdigest()


def _compile(src, verbose):
    path = osp.join(CSRC, src)
    obj = osp.join(OBJDIR, src[:-3] + ".o")
    stamp = obj + ".sha1"
    dig = _digest(path)
    if osp.exists(obj) and osp.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + ARCH + COMMON + EXTRA.get(src, []) + ["-c", path, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        print(r.stderr, flush=True)
    open(stamp, "w").write(dig)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(osp.join(OBJDIR, f))
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), sources()))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not osp.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
