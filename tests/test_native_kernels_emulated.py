"""CPU: the nine pointnet2 operator KERNELS (csrc/pointnet2_ops.cu, unmodified, launchers included; the bulk-TMA staging
of ball_query degraded to a plain copy) executed by the host emulator against the C oracle (oracle/pointnet2_ref.c, pinned
on the GPU by the unmodified reference kernels): indices bit-exact, including the adversarial tie / skip cases of
test_native_ops_gpu.py and the duplicated points of the reference's recorded demo sequence.  A no-GPU rehearsal of those
GPU tests' kernels -- same code, host execution model."""
import ctypes
import os.path as osp
import subprocess

import numpy as np
import pytest
import torch

from oracle.pointnet2_ref import RefExt
from pose2room_b200 import _lib, synthetic

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
NAMES = ("p2r_furthest_point_sampling", "p2r_ball_query", "p2r_gather_points", "p2r_gather_points_grad", "p2r_group_points",
         "p2r_group_points_grad", "p2r_three_nn", "p2r_three_interpolate", "p2r_three_interpolate_grad")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu_native") / "kernels_emu.so")
    inc = [osp.join(ROOT, "include"), osp.join(ROOT, "pose2room_b200", "csrc"), osp.join(ROOT, "tests", "csrc")]
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-std=c++17", "-w",
                    "-DP2R_HOST_EMULATION"] + sum((["-I", i] for i in inc), []) +
                   [osp.join(ROOT, "tests", "csrc", "kernels_emu.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    for name in NAMES:
        fn = getattr(lib, name)
        fn.argtypes = _lib.SIGNATURES[name]
        fn.restype = ctypes.c_int
    lib.emu_last_error.restype = ctypes.c_char_p
    return lib


def _p(a):
    return a.ctypes.data


def fps(lib, xyz, m):
    xyz = np.ascontiguousarray(xyz, np.float32)
    b, n, _ = xyz.shape
    out = np.full((b, m), -9, np.int32)
    scratch = np.zeros((b, n), np.float32)
    assert lib.p2r_furthest_point_sampling(_p(xyz), b, n, m, _p(out), _p(scratch), None) == 0, lib.emu_last_error()
    return torch.from_numpy(out)


def ball_query(lib, new_xyz, xyz, radius, nsample):
    new_xyz, xyz = np.ascontiguousarray(new_xyz, np.float32), np.ascontiguousarray(xyz, np.float32)
    b, m, n = new_xyz.shape[0], new_xyz.shape[1], xyz.shape[1]
    out = np.full((b, m, nsample), -9, np.int32)
    assert lib.p2r_ball_query(_p(new_xyz), _p(xyz), b, n, m, float(radius), nsample, _p(out), None) == 0, lib.emu_last_error()
    return torch.from_numpy(out)


@pytest.mark.parametrize("B,N,M", [(2, 512, 128), (3, 97, 13), (2, 1, 1), (2, 2, 2), (1, 768, 64), (1, 300, 300), (1, 1500, 40)])
def test_fps_kernel_under_emulation_index_exact(emu, B, N, M):
    xyz = synthetic.make_cloud(B, N, seed=N)
    assert torch.equal(fps(emu, xyz, M), RefExt.furthest_point_sampling(torch.from_numpy(xyz), M))


def test_fps_kernel_under_emulation_adversarial_ties_and_skips(emu):
    rng = np.random.default_rng(0)
    a = rng.integers(-2, 3, size=(2, 640, 3)).astype(np.float32) * 0.5       # lattice: masses of exact ties
    b = rng.normal(size=(2, 512, 3)).astype(np.float32)
    b[:, 100:140] = b[:, 0:40]                                                # duplicates
    b[:, 200:230] *= 0.01                                                     # |p|^2 <= 1e-3: skipped
    c = np.full((1, 768, 3), 0.5, np.float32)
    c[0, 100] = c[0, 600] = c[0, 7] = c[0, 519] = [3.0, 0.5, 0.5]
    for pts in (a, b, np.zeros((2, 64, 3), np.float32), c):
        m = min(48, pts.shape[1])
        assert torch.equal(fps(emu, pts, m), RefExt.furthest_point_sampling(torch.from_numpy(pts), m))


def test_fps_and_ball_query_kernels_on_the_recorded_demo_sequence(emu):
    """The hip trajectory of the reference's demo input resampled to 768 frames: 427 exactly repeated points."""
    g = np.load(osp.join(ROOT, "tests", "golden", "demo.npz"))
    hip = np.ascontiguousarray(g["raw_joints"][g["frame_ids"]][None, :, 0, :], np.float32)      # (1, 768, 3)
    assert len(np.unique(hip[0], axis=0)) < 400
    t = torch.from_numpy(hip)
    inds = fps(emu, hip, 128)
    assert torch.equal(inds, RefExt.furthest_point_sampling(t, 128))
    new_xyz = torch.gather(t, 1, inds.long()[:, :, None].expand(-1, -1, 3)).contiguous()
    for radius, ns in ((0.3, 16), (0.05, 4), (2.0, 64)):
        assert torch.equal(ball_query(emu, new_xyz.numpy(), hip, radius, ns), RefExt.ball_query(new_xyz, t, radius, ns))


@pytest.mark.parametrize("B,N,M,ns,r", [(2, 512, 128, 16, 0.3), (1, 5000, 33, 64, 0.2), (2, 100, 7, 8, 1e-4)])
def test_ball_query_kernel_under_emulation_index_exact(emu, B, N, M, ns, r):
    xyz = synthetic.make_cloud(B, N, seed=N + 1)
    new_xyz = np.ascontiguousarray(xyz[:, ::max(1, N // M)][:, :M])
    want = RefExt.ball_query(torch.from_numpy(new_xyz), torch.from_numpy(xyz), r, ns)
    assert torch.equal(ball_query(emu, new_xyz, xyz, r, ns), want)


def test_gather_group_interpolate_kernels_under_emulation(emu):
    rng = np.random.default_rng(3)
    B, C, N, M, S = 2, 19, 70, 23, 5
    pts = rng.normal(size=(B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, size=(B, M)).astype(np.int32)
    out = np.full((B, C, M), np.nan, np.float32)
    assert emu.p2r_gather_points(_p(pts), _p(idx), B, C, N, M, _p(out), None) == 0
    assert torch.equal(torch.from_numpy(out), RefExt.gather_points(torch.from_numpy(pts), torch.from_numpy(idx)))
    go = rng.normal(size=(B, C, M)).astype(np.float32)
    gp = np.zeros((B, C, N), np.float32)
    assert emu.p2r_gather_points_grad(_p(go), _p(idx), B, C, N, M, _p(gp), None) == 0
    assert torch.allclose(torch.from_numpy(gp), RefExt.gather_points_grad(torch.from_numpy(go), torch.from_numpy(idx), N), atol=1e-6)
    gidx = rng.integers(0, N, size=(B, M, S)).astype(np.int32)
    grouped = np.full((B, C, M, S), np.nan, np.float32)
    assert emu.p2r_group_points(_p(pts), _p(gidx), B, C, N, M, S, _p(grouped), None) == 0
    assert torch.equal(torch.from_numpy(grouped), RefExt.group_points(torch.from_numpy(pts), torch.from_numpy(gidx)))
    gg = rng.normal(size=(B, C, M, S)).astype(np.float32)
    gpp = np.zeros((B, C, N), np.float32)
    assert emu.p2r_group_points_grad(_p(gg), _p(gidx), B, C, N, M, S, _p(gpp), None) == 0
    assert torch.allclose(torch.from_numpy(gpp), RefExt.group_points_grad(torch.from_numpy(gg), torch.from_numpy(gidx), N), atol=1e-5)
    unknown, known = rng.normal(size=(B, 41, 3)).astype(np.float32), rng.normal(size=(B, M, 3)).astype(np.float32)
    d2, i3 = np.full((B, 41, 3), np.nan, np.float32), np.full((B, 41, 3), -9, np.int32)
    assert emu.p2r_three_nn(_p(unknown), _p(known), B, 41, M, _p(d2), _p(i3), None) == 0
    wd, wi = RefExt.three_nn(torch.from_numpy(unknown), torch.from_numpy(known))
    assert torch.equal(torch.from_numpy(i3), wi) and torch.equal(torch.from_numpy(d2), wd)
    feats, w = rng.normal(size=(B, C, M)).astype(np.float32), rng.random(size=(B, 41, 3)).astype(np.float32)
    interp = np.full((B, C, 41), np.nan, np.float32)
    assert emu.p2r_three_interpolate(_p(feats), _p(i3), _p(w), B, C, M, 41, _p(interp), None) == 0
    want = RefExt.three_interpolate(torch.from_numpy(feats), torch.from_numpy(i3), torch.from_numpy(w))
    assert torch.allclose(torch.from_numpy(interp), want, atol=1e-6)
    gi = rng.normal(size=(B, C, 41)).astype(np.float32)
    gf = np.zeros((B, C, M), np.float32)
    assert emu.p2r_three_interpolate_grad(_p(gi), _p(i3), _p(w), B, C, 41, M, _p(gf), None) == 0
    want = RefExt.three_interpolate_grad(torch.from_numpy(gi), torch.from_numpy(i3), torch.from_numpy(w), M)
    assert torch.allclose(torch.from_numpy(gf), want, atol=1e-5)
