"""CPU: the block-sparsity tables the graph-convolution GEMMs consume (k-block lists, tile masks, tile lists) against a
dense restatement, on the real 25- and 53-joint adjacency stacks of the reference's Graph (stgcn_layers.py:69-208)."""
import numpy as np
import pytest
import torch

from pose2room_b200.gemm_sm100 import BlockSparsity
from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency


def _pattern(joints):
    A = np.array(spatial_adjacency(layout_for_joints(joints), max_hop=5))
    return A, (np.abs(A).sum(0) > 0).T          # nz[w, v]: W_eff block (w, v) may be non-zero


@pytest.mark.parametrize("joints", [25, 53])
def test_pattern_matches_effective_weight(joints):
    """A block of W_eff = sum_k W_k (x) A_k is non-zero exactly where the pattern says (random W_k)."""
    A, nz = _pattern(joints)
    rng = np.random.default_rng(0)
    wk = rng.standard_normal((A.shape[0], 4, 4))                  # 4x4 channel blocks are enough for the structure
    w_eff = np.einsum("koi,kvw->wovi", wk, A).reshape(joints * 4, joints * 4)
    blocks = np.abs(w_eff).reshape(joints, 4, joints, 4).sum((1, 3)) > 0
    assert (blocks == nz).all()
    # every (v, w) pair belongs to at most one partition: the gradient-fold kernels rely on a few k per block only
    assert ((A != 0).sum(0) <= 1).all()


@pytest.mark.parametrize("block_n", [64, 128, 160, 256])
@pytest.mark.parametrize("transposed", [False, True])
def test_kb_lists_cover_exactly_the_support(block_n, transposed):
    _, nz = _pattern(25)
    sp = BlockSparsity(nz)
    tab = sp.kb_list(block_n, transposed, "cpu").numpy()
    pat = nz.T if transposed else nz                               # [n blocks, k blocks]
    n_cols = pat.shape[0] * 64
    assert tab.shape == (-(-n_cols // block_n), 1 + pat.shape[1])
    for i, row in enumerate(tab):
        ks = list(row[1:1 + row[0]])
        assert ks == sorted(set(ks))                               # ascending, no duplicates
        b0, b1 = (i * block_n) // 64, (min(n_cols, (i + 1) * block_n) - 1) // 64
        want = np.nonzero(pat[b0:b1 + 1].any(0))[0]
        assert ks == list(want)
    # skipping the unlisted k-blocks does not change a product with a weight that has this support
    rng = np.random.default_rng(1)
    w = rng.standard_normal((25 * 8, 25 * 8)) * np.kron(nz, np.ones((8, 8)))      # 8x8 stand-in blocks
    if transposed:
        w = w.T
    x = rng.standard_normal((5, 25 * 8))
    full = x @ w.T
    got = np.zeros_like(full)
    scale = 8 / 64                                                 # stand-in block size / real block size
    for i, row in enumerate(tab):
        c0, c1 = int(i * block_n * scale), int(min(25 * 64, (i + 1) * block_n) * scale)
        for kb in row[1:1 + row[0]]:
            got[:, c0:c1] += x[:, kb * 8:(kb + 1) * 8] @ w[c0:c1, kb * 8:(kb + 1) * 8].T
    assert np.allclose(got, full)


@pytest.mark.parametrize("bm,bn", [(128, 128), (256, 256), (128, 64)])
def test_tile_mask_and_tile_list(bm, bn):
    _, nz = _pattern(25)
    sp = BlockSparsity(nz)
    mask = sp.tile_mask(bm, bn, "cpu").numpy()
    tiles = sp.tile_list(bm, bn, "cpu")
    assert tiles.dtype == torch.int32 and tiles.shape[1] == 2
    assert sorted(map(tuple, tiles.numpy().tolist())) == sorted(map(tuple, np.argwhere(mask > 0).tolist()))
    keep = np.kron(mask, np.ones((bm, bn)))[:1600, :1600]
    assert (keep >= np.kron(nz, np.ones((64, 64)))).all()          # no structurally non-zero block is dropped
    assert 0.3 < sp.density < 0.7


def test_host_helpers_without_the_multi_stream_context():
    """Outside ops.overlap_weight_grads() the step helpers degrade to the plain behaviour (no streams, no workspace)."""
    from pose2room_b200 import ops
    assert ops.DEFER["on"] is False
    order = []
    out = ops.parallel_branches([lambda: order.append("a") or 1, lambda: order.append("b") or 2, lambda: order.append("c") or 3])
    assert out == [1, 2, 3] and order == ["a", "b", "c"]          # list order = issue order (RNG consumption order)
    z = ops.zeros_ws((3, 5), torch.float64, torch.device("cpu"))
    assert z.shape == (3, 5) and z.dtype == torch.float64 and float(z.abs().sum()) == 0.0


def test_internal_joint_order_is_a_permutation_that_tightens_the_tables():
    """stgcn._JOINT_ORDER_25 (throughput mode works in this joint order internally): a permutation of the 25 joints with
    fewer k-blocks per 256-wide n-tile and fewer non-zero 128x128 weight-gradient tiles than the reference order."""
    from pose2room_b200.p2rnet.stgcn import _JOINT_ORDER_25 as perm
    assert sorted(perm) == list(range(25))
    _, nz = _pattern(25)
    ident, mine = BlockSparsity(nz), BlockSparsity(nz[perm][:, perm])
    kb = lambda sp: int(sp.kb_list(256, False, "cpu")[:, 0].sum())
    tiles = lambda sp: int(sp.tile_mask(128, 128, "cpu").sum())
    assert (kb(ident), tiles(ident)) == (123, 109)
    assert (kb(mine), tiles(mine)) == (106, 99)
    assert mine.density == ident.density                           # the same blocks, moved
