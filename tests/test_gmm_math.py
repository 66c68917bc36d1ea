"""CPU: the arithmetic of the fused mixture-head kernels (csrc/gmm_math.h, compiled for the host by this test through
tests/csrc/gmm_mix_host.c, which mirrors the kernels' loops) against torch autograd over the module's own torch path
(pose2room_b200/p2rnet/mdn.py MixtureDensityHead.point_prediction = the reference's mdn.py:36-84, pinned end to end by
the reference goldens in test_model_oracle.py / test_model_gpu.py).  The CUDA kernels are covered by test_model_gpu.py."""
import ctypes
import os.path as osp
import subprocess

import numpy as np
import pytest
import torch

from pose2room_b200.p2rnet.mdn import MixtureDensityHead, Struct

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("gmm") / "gmm_mix_host.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", osp.join(ROOT, "pose2room_b200", "csrc"),
                    osp.join(ROOT, "tests", "csrc", "gmm_mix_host.c"), "-o", so, "-lm"], check=True)
    return ctypes.CDLL(so)


@pytest.mark.parametrize("G,D,mu_dtype", [(100, 3, torch.float32), (100, 2, torch.float64), (7, 3, torch.float32)])
def test_point_prediction_and_gradients_match_the_torch_path(host_lib, G, D, mu_dtype):
    rows = 37
    gen = torch.Generator().manual_seed(G * 10 + D)
    head = MixtureDensityHead(Struct(input_dim=8, num_gaussian=G, out_dim=D, n_samples=1, central_tendency="mean",
                                     mu_bias_init=torch.randn(G, D, generator=gen).to(mu_dtype)))
    with torch.no_grad():
        head.log_sigma.copy_(0.5 * torch.randn(G, D, generator=gen) - 0.5)
    logits = (2.0 * torch.randn(rows, G, generator=gen) - 1.0).requires_grad_(True)
    # the torch path, with the eps it draws made visible
    torch.manual_seed(5)
    out = head.point_prediction(torch.sigmoid(logits))
    torch.manual_seed(5)
    eps = head.mu.data.new(rows, G, 1, D).normal_()
    assert out.dtype == mu_dtype
    dout = torch.randn(rows, D, generator=gen).to(mu_dtype)
    out.backward(dout)

    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lg = logits.detach().numpy()
    mu = head.mu.detach().double().numpy().copy()
    ls = head.log_sigma.detach().numpy().copy()
    ep = eps.double().numpy().copy()
    got = np.zeros((rows, D))
    host_lib.host_gmm_mix(ptr(lg), ptr(mu), ptr(ls), ptr(ep), ctypes.c_longlong(rows), G, D, ptr(got))
    tol = 2e-6          # the torch path rounds sigmoid (and, for the float32 heads, everything) to float32
    assert np.abs(got - out.detach().double().numpy()).max() <= tol * max(1.0, np.abs(got).max())
    dlog, dmu, dls = np.zeros((rows, G), np.float32), np.zeros((G, D)), np.zeros((G, D), np.float32)
    do = dout.double().numpy().copy()
    host_lib.host_gmm_mix_grad(ptr(lg), ptr(mu), ptr(ls), ptr(ep), ptr(do), ctypes.c_longlong(rows), G, D, ptr(dlog),
                               ptr(dmu), ptr(dls))
    for name, a, b in [("logits", dlog, logits.grad), ("mu", dmu, head.mu.grad), ("log_sigma", dls, head.log_sigma.grad)]:
        b = b.double().numpy()
        assert np.abs(a - b).max() <= 3e-6 * max(1e-3, np.abs(b).max()), (name, np.abs(a - b).max(), np.abs(b).max())
