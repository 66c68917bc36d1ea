"""CPU: the C restatement of the nine native ops (oracle/pointnet2_ref.c) -- semantics and quirks,
and the reference's own Python wrappers running on top of it (goldens from tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle.pointnet2_ref import RefExt, knn_ref
from pose2room_b200 import synthetic


def brute_ball_query(new_xyz, xyz, r, ns):
    B, M, _ = new_xyz.shape
    out = np.zeros((B, M, ns), np.int32)
    r2 = np.float32(r) * np.float32(r)
    for b in range(B):
        for j in range(M):
            d = xyz[b].astype(np.float64) - new_xyz[b, j].astype(np.float64)
            d2 = (d * d).sum(1)
            hits = np.nonzero(d2 < r2 * (1 - 1e-6))[0][:ns]  # margin: only unambiguous hits
            if len(hits):
                out[b, j, :] = hits[0]
                out[b, j, :len(hits)] = hits
    return out


def test_ball_query_matches_bruteforce_away_from_the_boundary():
    xyz = synthetic.make_cloud(2, 200, seed=1)
    new = xyz[:, ::7].copy()
    got = RefExt.ball_query(torch.from_numpy(new), torch.from_numpy(xyz), 0.35, 8).numpy()
    want = brute_ball_query(new, xyz, 0.35, 8)
    d = np.linalg.norm(xyz[:, None] - new[:, :, None], axis=-1)
    ambiguous = (np.abs(d - 0.35) < 1e-5).any(axis=-1)
    assert np.array_equal(got[~ambiguous], want[~ambiguous])


def test_ball_query_quirks():
    xyz = torch.tensor([[[0, 0, 0], [0.1, 0, 0], [5, 5, 5], [0.2, 0, 0], [0.3, 0, 0]]], dtype=torch.float32)
    new = torch.tensor([[[0, 0, 0], [9, 9, 9]]], dtype=torch.float32)
    idx = RefExt.ball_query(new, xyz, 0.3, 4).numpy()
    # strict <: the point at distance exactly r (fp32 0.3*0.3 vs 0.3f^2) -- 0.3f*0.3f == r2 so not inside
    assert idx[0, 0].tolist() == [0, 1, 3, 0]      # padded with the first hit
    assert idx[0, 1].tolist() == [0, 0, 0, 0]      # no hit -> zeros


def test_fps_semantics():
    xyz = torch.from_numpy(synthetic.make_cloud(3, 300, seed=2))
    idx = RefExt.furthest_point_sampling(xyz, 40).numpy()
    assert (idx[:, 0] == 0).all()
    for b in range(3):
        assert len(set(idx[b].tolist())) == 40
        pts = xyz[b].numpy().astype(np.float64)
        # greedy property: each pick maximises the distance to the already-picked set
        for j in range(1, 10):
            dmin = np.min(((pts[:, None] - pts[idx[b, :j]][None]) ** 2).sum(-1), axis=1)
            assert dmin[idx[b, j]] >= dmin.max() * (1 - 1e-5)


def test_fps_skips_near_origin_points_and_tie_order():
    pts = np.zeros((1, 8, 3), np.float32)
    pts[0, :, 0] = [1, 0.01, 2, 2, 0.02, 3, 3, 1.5]   # 0.01, 0.02: |p|^2 <= 1e-3 -> never selected
    idx = RefExt.furthest_point_sampling(torch.from_numpy(pts), 6).numpy()[0]
    assert 1 not in idx and 4 not in idx
    # ties (points 5 and 6 equal): the reference's halving tree lets the thread id with a 0 in the lowest
    # differing bit survive -> 6 (0b110) beats 5 (0b101)
    assert idx[0] == 0 and idx[1] == 6
    # all points skipped -> index 0 repeated
    z = torch.zeros(1, 16, 3)
    assert RefExt.furthest_point_sampling(z, 5).numpy().tolist() == [[0, 0, 0, 0, 0]]


def test_fps_tie_order_depends_on_reference_block_size():
    # n = 768 -> reference block size 512: point k competes as thread k % 512; among equal distances the
    # smallest bit-reversed thread id wins: k = 600 (thread 88 = 0b001011000) beats k = 100
    # (0b001100100) because bit 2 is the lowest differing bit and 88 has a 0 there.
    pts = np.full((1, 768, 3), 0.5, np.float32)
    pts[0, 100] = pts[0, 600] = [3.0, 0.5, 0.5]
    idx = RefExt.furthest_point_sampling(torch.from_numpy(pts), 2).numpy()[0]
    assert RefExt.opt_n_threads(768) == 512
    assert idx.tolist() == [0, 600]
    # same thread (k = 7 and k = 519 are both thread 7): the earlier k wins (strict > in the scan)
    pts = np.full((1, 768, 3), 0.5, np.float32)
    pts[0, 7] = pts[0, 519] = [3.0, 0.5, 0.5]
    assert RefExt.furthest_point_sampling(torch.from_numpy(pts), 2).numpy()[0].tolist() == [0, 7]
    # thread 4 (0b100) vs thread 3 (0b011): lowest differing bit is bit 0 -> thread 4 wins
    pts = np.full((1, 768, 3), 0.5, np.float32)
    pts[0, 3] = pts[0, 4] = [3.0, 0.5, 0.5]
    assert RefExt.furthest_point_sampling(torch.from_numpy(pts), 2).numpy()[0].tolist() == [0, 4]


def test_three_nn_and_interpolate():
    rng = np.random.default_rng(0)
    unknown = torch.from_numpy(rng.normal(size=(2, 50, 3)).astype(np.float32))
    known = torch.from_numpy(rng.normal(size=(2, 20, 3)).astype(np.float32))
    d2, idx = RefExt.three_nn(unknown, known)
    full = ((unknown[:, :, None] - known[:, None]) ** 2).sum(-1)
    want = torch.argsort(full, dim=-1, stable=True)[..., :3]
    assert torch.equal(idx.long(), want)
    assert torch.allclose(d2, torch.gather(full, 2, want), rtol=1e-5, atol=1e-7)
    # fewer than three known points: unused slots are +inf with index 0
    d2, idx = RefExt.three_nn(unknown, known[:, :2].contiguous())
    assert torch.isinf(d2[..., 2]).all() and (idx[..., 2] == 0).all()


def test_gather_group_and_grads_accumulate_duplicates():
    pts = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5)
    idx = torch.tensor([[0, 0, 4], [1, 2, 2]], dtype=torch.int32)
    out = RefExt.gather_points(pts, idx)
    assert torch.equal(out, torch.gather(pts, 2, idx.long()[:, None].expand(2, 3, 3)))
    g = RefExt.gather_points_grad(torch.ones(2, 3, 3), idx, 5)
    assert g[0, 0].tolist() == [2, 0, 0, 0, 1]
    gidx = torch.tensor([[[0, 1], [1, 1]], [[4, 4], [3, 0]]], dtype=torch.int32)
    grouped = RefExt.group_points(pts, gidx)
    assert grouped.shape == (2, 3, 2, 2) and grouped[1, 2, 0, 1] == pts[1, 2, 4]
    gg = RefExt.group_points_grad(torch.ones(2, 3, 2, 2), gidx, 5)
    assert gg[0, 1].tolist() == [1, 3, 0, 0, 0]


def test_empty_batch_is_fine():
    assert RefExt.furthest_point_sampling(torch.zeros(0, 4, 3), 2).shape == (0, 2)
    assert RefExt.ball_query(torch.zeros(0, 2, 3), torch.zeros(0, 4, 3), 0.3, 4).shape == (0, 2, 4)


def test_reference_wrappers_on_oracle_match_committed_goldens(golden_pointnet2):
    """The goldens were produced by the reference's pointnet2_utils / pointnet2_modules running over this
    oracle; re-deriving the native parts here guards the oracle against silent edits."""
    g = golden_pointnet2
    xyz = torch.from_numpy(g["xyz"])
    assert np.array_equal(RefExt.furthest_point_sampling(xyz, g["fps"].shape[1]).numpy(), g["fps"])
    new_xyz = torch.from_numpy(g["new_xyz"])
    idx = RefExt.ball_query(new_xyz, xyz, 0.4, 8)
    grouped = RefExt.group_points(torch.from_numpy(g["feats"]), idx)
    assert np.array_equal(grouped.numpy(), g["qg_features"][:, 3:])
    d2, tidx = RefExt.three_nn(xyz, new_xyz)
    assert np.array_equal(tidx.numpy(), g["tnn_idx"])
    assert np.array_equal(torch.sqrt(d2).numpy(), g["tnn_dist"])


def test_knn_oracle_matches_reference_golden(golden_pointnet2):
    g = golden_pointnet2
    idx = knn_ref(torch.from_numpy(g["knn_x"]), 8).numpy()
    # reference = torch.topk on a matmul-based distance: same neighbour SETS; order can differ on near-ties
    assert np.array_equal(np.sort(idx, -1), np.sort(g["knn_idx"], -1))
    assert (idx[..., 0] == np.arange(64)[None]).all()
