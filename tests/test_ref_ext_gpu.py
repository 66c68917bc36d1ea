"""GPU: the UNMODIFIED reference CUDA extension (compiled for sm_100a into oracle/_ref by
oracle/build_ref_ext.py) as referee: it pins the CPU oracle AND the new kernels on the same inputs."""
import numpy as np
import pytest
import torch

from oracle import build_ref_ext
from oracle.pointnet2_ref import RefExt
from pose2room_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(cuda):
    mod = build_ref_ext.load_ref_ext()
    if mod is None:
        pytest.skip("oracle/_ref/p2r_ref_ext.so not built (needs /root/reference at build time)")
    return mod


def adversarial_clouds():
    rng = np.random.default_rng(42)
    yield torch.from_numpy(synthetic.make_cloud(4, 512, seed=1))
    yield torch.from_numpy(synthetic.make_cloud(2, 768, seed=2))
    yield torch.from_numpy(synthetic.make_cloud(2, 3000, seed=3))
    yield torch.from_numpy(rng.integers(-2, 3, size=(3, 640, 3)).astype(np.float32) * 0.5)  # exact ties
    dup = rng.normal(size=(2, 512, 3)).astype(np.float32)
    dup[:, 256:] = dup[:, :256]
    dup[:, 10:20] *= 0.01
    yield torch.from_numpy(dup)


def test_fps_three_way(cuda, ref):
    from pose2room_b200 import ext
    for xyz in adversarial_clouds():
        m = min(128, xyz.shape[1])
        r = ref.furthest_point_sampling(xyz.to(cuda), m).cpu()
        assert torch.equal(RefExt.furthest_point_sampling(xyz, m), r), "CPU oracle != reference kernel"
        assert torch.equal(ext.furthest_point_sampling(xyz.to(cuda), m).cpu(), r), "new kernel != reference kernel"


def test_ball_query_three_way(cuda, ref):
    from pose2room_b200 import ext
    for xyz in adversarial_clouds():
        m = min(128, xyz.shape[1])
        sel = ref.furthest_point_sampling(xyz.to(cuda), m).long()
        new_xyz = torch.gather(xyz.to(cuda), 1, sel[:, :, None].expand(-1, -1, 3)).contiguous()
        for r_, ns in [(0.3, 16), (0.5, 64), (1.0, 8)]:
            r = ref.ball_query(new_xyz, xyz.to(cuda), r_, ns).cpu()
            assert torch.equal(RefExt.ball_query(new_xyz.cpu(), xyz, r_, ns), r)
            assert torch.equal(ext.ball_query(new_xyz, xyz.to(cuda), r_, ns).cpu(), r)


def test_three_nn_interpolate_group_gather_three_way(cuda, ref):
    from pose2room_b200 import ext
    g = torch.Generator().manual_seed(0)
    for xyz in adversarial_clouds():
        known = xyz[:, ::5].contiguous()
        d_r, i_r = ref.three_nn(xyz.to(cuda), known.to(cuda))
        d_o, i_o = RefExt.three_nn(xyz, known)
        d_n, i_n = ext.three_nn(xyz.to(cuda), known.to(cuda))
        assert torch.equal(i_o, i_r.cpu()) and torch.equal(d_o, d_r.cpu())
        assert torch.equal(i_n, i_r) and torch.equal(d_n, d_r)
        feats = torch.randn(xyz.shape[0], 19, known.shape[1], generator=g).to(cuda)
        w = torch.rand(xyz.shape[0], xyz.shape[1], 3, generator=g).to(cuda)
        assert torch.equal(ext.three_interpolate(feats, i_r, w), ref.three_interpolate(feats, i_r, w))
        assert torch.equal(RefExt.three_interpolate(feats.cpu(), i_r.cpu(), w.cpu()), ref.three_interpolate(feats, i_r, w).cpu())
        go = torch.randn(xyz.shape[0], 19, xyz.shape[1], generator=g).to(cuda)
        assert torch.allclose(ext.three_interpolate_grad(go, i_r, w, known.shape[1]),
                              ref.three_interpolate_grad(go, i_r, w, known.shape[1]), rtol=1e-4, atol=1e-5)
        idx = torch.randint(0, xyz.shape[1], (xyz.shape[0], 32, 16), generator=g, dtype=torch.int32).to(cuda)
        pts = torch.randn(xyz.shape[0], 24, xyz.shape[1], generator=g).to(cuda)
        assert torch.equal(ext.group_points(pts, idx), ref.group_points(pts, idx))
        gg = torch.randn(xyz.shape[0], 24, 32, 16, generator=g).to(cuda)
        assert torch.allclose(ext.group_points_grad(gg, idx, xyz.shape[1]), ref.group_points_grad(gg, idx, xyz.shape[1]),
                              rtol=1e-4, atol=1e-5)
        i1 = idx[:, :, 0].contiguous()
        assert torch.equal(ext.gather_points(pts, i1), ref.gather_points(pts, i1))
        go2 = torch.randn(xyz.shape[0], 24, 32, generator=g).to(cuda)
        assert torch.allclose(ext.gather_points_grad(go2, i1, xyz.shape[1]), ref.gather_points_grad(go2, i1, xyz.shape[1]),
                              rtol=1e-4, atol=1e-5)


def test_reference_sa_module_runs_unmodified_on_new_kernels(cuda, ref, golden_pointnet2):
    """Drop-in proof at the operator ABI: an `_ext`-shaped object is all the reference Python needs.
    Here the golden (reference Python over the CPU oracle) is reproduced by OUR pointnet2_modules port on GPU."""
    from pose2room_b200.pointnet2_modules import PointnetSAModuleVotes
    g = golden_pointnet2
    C = g["feats"].shape[1]
    sa = PointnetSAModuleVotes(npoint=g["sa_inds"].shape[1], radius=0.4, nsample=8, mlp=[C, 16, 12], use_xyz=False,
                               normalize_xyz=True, bn=False).to(cuda)
    with torch.no_grad():
        sa.mlp_module[0].weight.copy_(torch.from_numpy(g["sa_w0"]))
        sa.mlp_module[0].bias.copy_(torch.from_numpy(g["sa_b0"]))
        sa.mlp_module[2].weight.copy_(torch.from_numpy(g["sa_w1"]))
        sa.mlp_module[2].bias.copy_(torch.from_numpy(g["sa_b1"]))
    feats = torch.from_numpy(g["feats"]).to(cuda).requires_grad_(True)
    xyz, feat, inds = sa(torch.from_numpy(g["xyz"]).to(cuda), feats)
    assert np.array_equal(inds.cpu().numpy(), g["sa_inds"])
    assert np.array_equal(xyz.cpu().numpy(), g["sa_xyz"])
    assert np.allclose(feat.detach().cpu().numpy(), g["sa_feat"], rtol=1e-4, atol=1e-5)
    (feat * torch.from_numpy(g["sa_w"]).to(cuda)).sum().backward()
    assert np.allclose(feats.grad.cpu().numpy(), g["sa_feats_grad"], rtol=1e-4, atol=1e-5)
    assert np.allclose(sa.mlp_module[0].weight.grad.cpu().numpy(), g["sa_w0_grad"], rtol=1e-4, atol=1e-4)
