"""ctypes front-end of oracle/pointnet2_ref.c: an `_ext`-shaped object on CPU torch tensors.

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this.  `RefExt` presents the reference's native operator ABI
(`pointnet2_ops._ext`, _ext-src/src/bindings.cpp:6-19) so the UNMODIFIED reference Python
(pointnet2_utils.py / pointnet2_modules.py / ProposalNet) can run on CPU on top of it.
"""
import ctypes
import os
import os.path as osp
import subprocess

import numpy as np
import torch

_HERE = osp.dirname(osp.abspath(__file__))
_SRC = osp.join(_HERE, "pointnet2_ref.c")
_OUT = osp.join(_HERE, "_build", "libp2r_oracle.so")


def build(force=False):
    if force or not osp.exists(_OUT) or osp.getmtime(_OUT) < osp.getmtime(_SRC):
        os.makedirs(osp.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _OUT, _SRC, "-lm"])
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


class RefExt:
    """Same nine names / argument order as pointnet2_ops._ext (bindings.cpp:6-19)."""

    @staticmethod
    def opt_n_threads(n):
        return lib().p2r_ref_opt_n_threads(int(n))

    @staticmethod
    def furthest_point_sampling(points, nsamples):
        b, n, _ = points.shape
        out = torch.zeros(b, nsamples, dtype=torch.int32)
        lib().p2r_ref_furthest_point_sampling(b, n, int(nsamples), _f(points), _i(out))
        return out

    @staticmethod
    def gather_points(points, idx):
        b, c, n = points.shape
        m = idx.shape[1]
        out = torch.zeros(b, c, m, dtype=torch.float32)
        lib().p2r_ref_gather_points(b, c, n, m, _f(points), _i(idx), _f(out))
        return out

    @staticmethod
    def gather_points_grad(grad_out, idx, n):
        b, c, m = grad_out.shape
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().p2r_ref_gather_points_grad(b, c, int(n), m, _f(grad_out), _i(idx), _f(out))
        return out

    @staticmethod
    def ball_query(new_xyz, xyz, radius, nsample):
        b, m, _ = new_xyz.shape
        n = xyz.shape[1]
        out = torch.zeros(b, m, nsample, dtype=torch.int32)
        lib().p2r_ref_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _f(new_xyz), _f(xyz), _i(out))
        return out

    @staticmethod
    def group_points(points, idx):
        b, c, n = points.shape
        _, npoints, nsample = idx.shape
        out = torch.zeros(b, c, npoints, nsample, dtype=torch.float32)
        lib().p2r_ref_group_points(b, c, n, npoints, nsample, _f(points), _i(idx), _f(out))
        return out

    @staticmethod
    def group_points_grad(grad_out, idx, n):
        b, c, npoints, nsample = grad_out.shape
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().p2r_ref_group_points_grad(b, c, int(n), npoints, nsample, _f(grad_out), _i(idx), _f(out))
        return out

    @staticmethod
    def three_nn(unknown, known):
        b, n, _ = unknown.shape
        m = known.shape[1]
        dist2 = torch.zeros(b, n, 3, dtype=torch.float32)
        idx = torch.zeros(b, n, 3, dtype=torch.int32)
        lib().p2r_ref_three_nn(b, n, m, _f(unknown), _f(known), _f(dist2), _i(idx))
        return dist2, idx

    @staticmethod
    def three_interpolate(points, idx, weight):
        b, c, m = points.shape
        n = idx.shape[1]
        out = torch.zeros(b, c, n, dtype=torch.float32)
        lib().p2r_ref_three_interpolate(b, c, m, n, _f(points), _i(idx), _f(weight), _f(out))
        return out

    @staticmethod
    def three_interpolate_grad(grad_out, idx, weight, m):
        b, c, n = grad_out.shape
        out = torch.zeros(b, c, m, dtype=torch.float32)
        lib().p2r_ref_three_interpolate_grad(b, c, n, int(m), _f(grad_out), _i(idx), _f(weight), _f(out))
        return out


def knn_ref(x, k):
    """net_utils/vn_dgcnn_util.py:4-10 restated: x (B,C,N) -> idx (B,N,k) int64.
    pairwise = -xx - (-2 x^T x) - xx^T, top-k largest; ties resolved lowest index first
    (torch.topk's CPU order on equal values is unspecified; the reference never pins it)."""
    x = x.detach().cpu().float()
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = (-xx - inner - xx.transpose(2, 1)).numpy()
    n = pd.shape[-1]
    order = np.lexsort((np.broadcast_to(np.arange(n), pd.shape), -pd), axis=-1)
    return torch.from_numpy(order[..., :k].copy())
