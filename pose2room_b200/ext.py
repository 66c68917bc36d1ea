"""`pointnet2_ops._ext`-shaped operator ABI on top of libp2r_b200.so.

Mirrors the nine pybind entry points of the reference's native extension
(/root/reference/external/pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19): same
names, argument order, dtypes, shapes and error behaviour (CHECK_CONTIGUOUS / CHECK_IS_FLOAT /
CHECK_IS_INT raise RuntimeError, _ext-src/include/utils.h:5-25; non-CUDA tensors raise like the
reference's "CPU not supported" asserts).  Outputs are fresh base tensors (never views):
QueryAndGroup edits the grouping result in place (pointnet2_utils.py:335,337).

`install_as_pointnet2_ops()` publishes this module as `pointnet2_ops._ext` so the UNMODIFIED
reference Python (pointnet2_utils.py:7-8) picks these kernels up with zero edits.
"""
import sys
import types

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, name, dtype):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (CPU not supported)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be a%s tensor" % (name, " float" if dtype == torch.float32 else "n int"))
    return t.data_ptr()


def furthest_point_sampling(points, nsamples):
    p = _check(points, "points", torch.float32)
    b, n, _ = points.shape
    out = torch.zeros(b, nsamples, dtype=torch.int32, device=points.device)
    scratch = torch.empty(b * n, dtype=torch.float32, device=points.device) if n > 32768 else None
    with torch.cuda.device(points.device):
        _lib.call("p2r_furthest_point_sampling", p, b, n, int(nsamples), out.data_ptr(),
                  scratch.data_ptr() if scratch is not None else None, _stream())
    return out


def gather_points(points, idx):
    p = _check(points, "points", torch.float32)
    i = _check(idx, "idx", torch.int32)
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty(b, c, m, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.call("p2r_gather_points", p, i, b, c, n, m, out.data_ptr(), _stream())
    return out


def gather_points_grad(grad_out, idx, n):
    g = _check(grad_out, "grad_out", torch.float32)
    i = _check(idx, "idx", torch.int32)
    b, c, m = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.call("p2r_gather_points_grad", g, i, b, c, int(n), m, out.data_ptr(), _stream())
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    q = _check(new_xyz, "new_xyz", torch.float32)
    p = _check(xyz, "xyz", torch.float32)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.empty(b, m, nsample, dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.call("p2r_ball_query", q, p, b, n, m, float(radius), int(nsample), out.data_ptr(), _stream())
    return out


def group_points(points, idx):
    p = _check(points, "points", torch.float32)
    i = _check(idx, "idx", torch.int32)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty(b, c, npoints, nsample, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.call("p2r_group_points", p, i, b, c, n, npoints, nsample, out.data_ptr(), _stream())
    return out


def group_points_grad(grad_out, idx, n):
    g = _check(grad_out, "grad_out", torch.float32)
    i = _check(idx, "idx", torch.int32)
    b, c, npoints, nsample = grad_out.shape
    out = torch.zeros(b, c, n, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.call("p2r_group_points_grad", g, i, b, c, int(n), npoints, nsample, out.data_ptr(), _stream())
    return out


def three_nn(unknown, known):
    u = _check(unknown, "unknowns", torch.float32)
    k = _check(known, "knows", torch.float32)
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.empty(b, n, 3, dtype=torch.float32, device=unknown.device)
    idx = torch.empty(b, n, 3, dtype=torch.int32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        _lib.call("p2r_three_nn", u, k, b, n, m, dist2.data_ptr(), idx.data_ptr(), _stream())
    return dist2, idx


def three_interpolate(points, idx, weight):
    p = _check(points, "points", torch.float32)
    i = _check(idx, "idx", torch.int32)
    w = _check(weight, "weight", torch.float32)
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty(b, c, n, dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        _lib.call("p2r_three_interpolate", p, i, w, b, c, m, n, out.data_ptr(), _stream())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    g = _check(grad_out, "grad_out", torch.float32)
    i = _check(idx, "idx", torch.int32)
    w = _check(weight, "weight", torch.float32)
    b, c, n = grad_out.shape
    out = torch.zeros(b, c, m, dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        _lib.call("p2r_three_interpolate_grad", g, i, w, b, c, n, int(m), out.data_ptr(), _stream())
    return out


NAMES = ["furthest_point_sampling", "gather_points", "gather_points_grad", "three_nn", "three_interpolate",
         "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"]


def install_as_pointnet2_ops(package_path=None):
    """Register this module as `pointnet2_ops._ext` (and a `pointnet2_ops` package whose __path__ points
    at `package_path`, e.g. <reference>/external/pointnet2_ops_lib/pointnet2_ops) BEFORE the reference's
    pointnet2_utils is imported.  See INTEGRATION.md."""
    _lib.load()
    me = sys.modules[__name__]
    pkg = sys.modules.get("pointnet2_ops")
    if pkg is None:
        pkg = types.ModuleType("pointnet2_ops")
        pkg.__path__ = [package_path] if package_path else []
        sys.modules["pointnet2_ops"] = pkg
    pkg._ext = me
    sys.modules["pointnet2_ops._ext"] = me
    return me
