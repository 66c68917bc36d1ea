"""GPU: the tcgen05/TMA bf16 GEMM against torch.matmul (fp32 accumulate reference on the same bf16 inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def _check(c, ref, tol=2e-2):
    err = (c.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale + 1e-3, (err, scale)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (256, 128, 128, 128), (384, 256, 192, 256), (1000, 200, 320, 0),
                                       (512, 1600, 1600, 0), (4096, 64, 64, 0), (130, 72, 40, 0)])
def test_gemm_layouts_and_shapes(cuda, M, N, K, bn, a_mn, b_mn):
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn((K, M) if a_mn else (M, K), generator=g).to(cuda).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), generator=g).to(cuda).bfloat16()
    if (a_mn and M % 8) or (b_mn and N % 8) or ((not a_mn or not b_mn) and K % 8):
        pytest.skip("pitch not a multiple of 16 bytes for this layout")
    ref = _ref(a, b, a_mn, b_mn)
    for out_dtype in [torch.float32, torch.bfloat16]:
        c = gemm_sm100.gemm(a, b, a_mn, b_mn, out_dtype=out_dtype, block_n=bn)
        _check(c, ref, tol=1e-2 if out_dtype == torch.bfloat16 else 1e-5 * K ** 0.5 + 1e-4)


def test_gemm_exact_on_small_integers(cuda):
    """Integer-valued bf16 operands: every product and partial sum is exact in fp32, so the result must be EXACT --
    this catches any descriptor / swizzle / k-slice mix-up that random data could hide inside a tolerance."""
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(0)
    for (M, N, K, a_mn, b_mn, bn) in [(256, 160, 256, False, False, 160), (256, 128, 256, False, True, 128),
                                      (256, 128, 512, True, True, 128), (128, 256, 128, True, False, 256),
                                      (384, 320, 192, False, False, 160), (256, 64, 192, False, False, 64)]:
        a = torch.randint(-4, 5, (K, M) if a_mn else (M, K), generator=g).float().to(cuda).bfloat16()
        b = torch.randint(-4, 5, (K, N) if b_mn else (N, K), generator=g).float().to(cuda).bfloat16()
        c = gemm_sm100.gemm(a, b, a_mn, b_mn, out_dtype=torch.float32, block_n=bn)
        assert torch.equal(c, _ref(a, b, a_mn, b_mn)), (M, N, K, a_mn, b_mn, bn)


def test_gemm_epilogue_bias_relu_and_splitk(cuda):
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(1)
    a = torch.randn(700, 256, generator=g).to(cuda).bfloat16()
    b = (torch.randn(259 - 3, 256, generator=g) / 16).to(cuda).bfloat16()
    bias = torch.randn(256, generator=g).to(cuda)
    ref = torch.relu(a.float() @ b.float().t() + bias)
    _check(gemm_sm100.gemm(a, b, bias=bias, relu=True, out_dtype=torch.float32), ref, tol=1e-4)
    _check(gemm_sm100.gemm(a, b, bias=bias, relu=True, out_dtype=torch.bfloat16), ref, tol=1e-2)
    # weight-gradient shape: huge reduction, small output, split-K with fp32 atomics
    dz = torch.randn(20000, 64, generator=g).to(cuda).bfloat16()
    x = torch.randn(20000, 192, generator=g).to(cuda).bfloat16()
    ref = dz.float().t() @ x.float()
    _check(gemm_sm100.gemm(dz, x, True, True, out_dtype=torch.float32, splits=5), ref, tol=1e-4)
    _check(gemm_sm100.gemm(dz, x, True, True, out_dtype=torch.float32, splits=1), ref, tol=1e-4)


def test_linear_autograd_bf16_backend(cuda):
    """ops.linear with the tensor-core backend installed: forward, dx, dW, db vs an fp32 torch reference on the
    same (bf16-rounded) inputs."""
    from pose2room_b200 import gemm_sm100, ops
    gemm_sm100.install()
    try:
        g = torch.Generator().manual_seed(2)
        for (M, K, N) in [(2048, 1600, 1600), (8192, 64, 64), (4096, 192, 64), (1024, 256, 256), (4096, 256, 259),
                          (4096, 128, 100)]:   # the last two: odd widths, zero-padded to a multiple of 8
            x = torch.randn(M, K, generator=g).to(cuda).bfloat16().requires_grad_(True)
            w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda).requires_grad_(True)
            b = torch.randn(N, generator=g).to(cuda).requires_grad_(True)
            y = ops.linear(x, w, b, relu=True)
            assert y.dtype == torch.bfloat16
            xr = x.detach().float().requires_grad_(True)
            wr = w.detach().bfloat16().float().requires_grad_(True)
            br = b.detach().clone().requires_grad_(True)
            yr = torch.relu(xr @ wr.t() + br)
            _check(y, yr.detach(), tol=1e-2)
            go = torch.randn(M, N, generator=g).to(cuda).bfloat16()
            gx, gw, gb = torch.autograd.grad(y, [x, w, b], go)
            # reference backward with the SAME relu mask as the bf16 forward
            mask = (y.detach().float() > 0).float()
            dz = go.float() * mask
            _check(gx, dz @ wr.detach(), tol=1e-2)
            _check(gw, dz.t() @ xr.detach(), tol=1e-3)
            _check(gb, dz.sum(0), tol=1e-3)
    finally:
        gemm_sm100.uninstall()


def test_implicit_temporal_conv_fwd_dx_dw(cuda):
    """The 3-tap temporal conv as implicit GEMMs (3-D TMA maps, zero fill at sequence ends) vs nn.Conv2d in fp32 on the
    same bf16-rounded operands; integer-valued data makes the forward and the input gradient exact."""
    import torch.nn as nn
    from pose2room_b200 import gemm_sm100, ops
    gemm_sm100.install()
    try:
        g = torch.Generator().manual_seed(3)
        for (B, T, V) in [(2, 128, 25), (3, 64, 26), (1, 1024, 25)]:
            C = 64
            conv = nn.Conv2d(C, C, (3, 1), (1, 1), (1, 0)).to(cuda)
            with torch.no_grad():
                conv.weight.copy_(torch.randint(-2, 3, conv.weight.shape, generator=g).float())
                conv.bias.copy_(torch.randint(-2, 3, (C,), generator=g).float())
            x = torch.randint(-2, 3, (B, T, V, C), generator=g).float().to(cuda).bfloat16().requires_grad_(True)
            assert gemm_sm100._Backend.supports_tconv(x.shape, C)
            y = ops.temporal_conv(x, conv.weight, conv.bias)
            xr = x.detach().float().requires_grad_(True)
            ref = conv(xr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).reshape(B * T * V, C)
            assert torch.equal(y.float(), ref.detach()), (B, T, V)
            go = torch.randint(-2, 3, (B * T * V, C), generator=g).float().to(cuda)
            gx, gw, gb = torch.autograd.grad(y, [x, conv.weight, conv.bias], go.bfloat16())
            rx, rw, rb = torch.autograd.grad(ref, [xr, conv.weight, conv.bias], go)
            assert torch.equal(gx.float(), rx), (B, T, V)
            assert torch.allclose(gw, rw, rtol=1e-5, atol=1e-2) and torch.allclose(gb, rb, rtol=1e-5, atol=1e-2)
    finally:
        gemm_sm100.uninstall()


def _random_block_pattern(nb, kb, g, density=0.5):
    nz = torch.rand(nb, kb, generator=g) < density
    nz |= torch.eye(nb, kb, dtype=torch.bool)          # every n-block keeps at least one k-block
    return nz.numpy()


@pytest.mark.parametrize("bn", [64, 128, 160, 256])
def test_gemm_block_sparse_reduction_exact(cuda, bn):
    """kb_list: only the listed 64-wide k-blocks of each n-tile are loaded / multiplied.  With a weight that IS zero in
    the skipped blocks the result must equal the dense product -- exactly, on integer-valued operands."""
    import numpy as np
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(10 + bn)
    M, NB, KB = 384, 10, 10                        # N = K = 640 (a multiple of 160 and of 64)
    nz = _random_block_pattern(NB, KB, g)
    sp = gemm_sm100.BlockSparsity(nz)
    w = torch.randint(-3, 4, (NB * 64, KB * 64), generator=g).float()
    w *= torch.from_numpy(np.kron(nz, np.ones((64, 64), dtype=np.float32)))
    x = torch.randint(-3, 4, (M, KB * 64), generator=g).float()
    xb, wb = x.to(cuda).bfloat16(), w.to(cuda).bfloat16()
    c = gemm_sm100.gemm(xb, wb, out_dtype=torch.float32, block_n=bn, kb_list=sp.kb_list(bn, False, cuda))
    assert torch.equal(c.cpu(), x @ w.t())
    # input-gradient orientation: dx = dz . W, reduction over W's rows, pattern transposed
    dz = torch.randint(-3, 4, (M, NB * 64), generator=g).float()
    c = gemm_sm100.gemm(dz.to(cuda).bfloat16(), wb.t().contiguous(), out_dtype=torch.float32, block_n=bn,
                        kb_list=sp.kb_list(bn, True, cuda))
    assert torch.equal(c.cpu(), dz @ w)
    # a k-block list shorter than the weight's support really skips work: drop everything -> bias only
    empty = torch.zeros_like(sp.kb_list(bn, False, cuda))
    bias = torch.arange(NB * 64, dtype=torch.float32, device=cuda)
    c = gemm_sm100.gemm(xb, wb, bias=bias, out_dtype=torch.float32, block_n=bn, kb_list=empty)
    assert torch.equal(c, bias[None].expand(M, -1))


def test_gemm_tile_mask_weight_gradient(cuda):
    """tile_mask: structurally-zero 128x128 tiles of dW = dz^T.x are skipped and stay zero, the others are exact."""
    import numpy as np
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(11)
    M, NB, KB = 1024, 6, 6
    nz = _random_block_pattern(NB, KB, g, density=0.3)
    sp = gemm_sm100.BlockSparsity(nz)
    dz = torch.randint(-2, 3, (M, NB * 64), generator=g).float()
    x = torch.randint(-2, 3, (M, KB * 64), generator=g).float()
    mask = sp.tile_mask(128, 128, cuda)
    c = gemm_sm100.gemm(dz.to(cuda).bfloat16(), x.to(cuda).bfloat16(), True, True, out_dtype=torch.float32, block_n=128,
                        tile_mask=mask)
    full = dz.t() @ x
    keep = torch.from_numpy(np.kron(mask.cpu().numpy(), np.ones((128, 128), dtype=np.float32)))[:NB * 64, :KB * 64]
    assert torch.equal(c.cpu(), full * keep)
    # every block the pattern marks non-zero lies inside a kept tile
    assert (keep.numpy() >= np.kron(nz, np.ones((64, 64), dtype=np.float32))).all()


@pytest.mark.parametrize("M,N,K,bn", [(384, 640, 256, 160), (1000, 128, 192, 128), (4096, 64, 64, 64), (130, 320, 64, 160),
                                       (33000, 64, 192, 64)])
def test_gemm_fused_output_statistics(cuda, M, N, K, bn):
    """stats: per-channel (column % 64) sum and sum of squares of the STORED bf16 output, from the GEMM epilogue."""
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).to(cuda).bfloat16()
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda)
    for copies in (1, 16):
        stats = torch.zeros(copies, 2, 64, dtype=torch.float64, device=cuda)
        c = gemm_sm100.gemm(a, b, bias=bias, out_dtype=torch.bfloat16, block_n=bn, stats=stats)
        ref = gemm_sm100.gemm(a, b, bias=bias, out_dtype=torch.bfloat16, block_n=bn)
        assert torch.equal(c, ref)                                      # the output itself is unchanged
        cd = c.double().reshape(M, N // 64, 64)
        s = stats.sum(0)
        assert torch.allclose(s[0], cd.sum((0, 1)), rtol=1e-5, atol=1e-3 * M ** 0.5), (s[0] - cd.sum((0, 1))).abs().max()
        assert torch.allclose(s[1], (cd * cd).sum((0, 1)), rtol=1e-5, atol=1e-3), (s[1] - (cd * cd).sum((0, 1))).abs().max()


def test_tconv_and_batchnorm_with_fused_statistics(cuda):
    """temporal_conv(..., want_stats=True) + batchnorm_act(..., sums=...) == the two-pass BatchNorm on the same y."""
    import torch.nn as nn
    from pose2room_b200 import gemm_sm100, ops
    gemm_sm100.install()
    try:
        g = torch.Generator().manual_seed(12)
        B, T, V, C = 2, 256, 25, 64
        conv = nn.Conv2d(C, C, (3, 1), (1, 1), (1, 0)).to(cuda)
        x = torch.randn(B, T, V, C, generator=g).to(cuda).bfloat16()
        y, sums = ops.temporal_conv(x, conv.weight, conv.bias, want_stats=True)
        assert sums.numel() > 0 and sums.shape[1:] == (2, 64)
        y2 = ops.temporal_conv(x, conv.weight, conv.bias)
        assert torch.equal(y, y2)
        yd = y.double()
        assert torch.allclose(sums.sum(0)[0], yd.sum(0), rtol=1e-6, atol=1e-2)
        assert torch.allclose(sums.sum(0)[1], (yd * yd).sum(0), rtol=1e-6, atol=1e-2)
        bn_a, bn_b = nn.BatchNorm2d(C).to(cuda).train(), nn.BatchNorm2d(C).to(cuda).train()
        out_a = ops.batchnorm_act(y, bn_a, relu=True, sums=sums)
        out_b = ops.batchnorm_act(y, bn_b, relu=True)
        assert torch.allclose(out_a.float(), out_b.float(), atol=2e-2)
        assert torch.allclose(bn_a.running_mean, bn_b.running_mean, atol=1e-6)
        assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-6, atol=1e-7)
    finally:
        gemm_sm100.uninstall()


def test_graph_conv_linear_sparse_matches_dense(cuda):
    """ops.linear on a real W_eff (25 joints, max_hop 5) with the block pattern from the adjacency: forward, dx and dW
    equal the dense run (the skipped blocks are exact zeros of W_eff; dW is compared on the pattern's support)."""
    import numpy as np
    from pose2room_b200 import gemm_sm100, ops
    from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency
    gemm_sm100.install()
    try:
        g = torch.Generator().manual_seed(13)
        A = torch.tensor(np.array(spatial_adjacency(layout_for_joints(25), max_hop=5)), dtype=torch.float32)
        sp = gemm_sm100.BlockSparsity((A.abs().sum(0) > 0).t().numpy())
        assert 0.4 < sp.density < 0.6
        wk = torch.randn(11, 64, 64, generator=g) / 8
        w_eff = torch.einsum("koi,kvw->wovi", wk, A).reshape(1600, 1600).to(cuda)
        support = torch.from_numpy(np.kron(sp.nz, np.ones((64, 64), dtype=np.float32))).to(cuda)
        assert torch.equal(w_eff * support, w_eff)
        bias = torch.randn(1600, generator=g).to(cuda)
        outs = []
        for s in (None, sp):
            x = torch.randn(1024, 1600, generator=torch.Generator().manual_seed(99)).to(cuda).bfloat16().requires_grad_(True)
            w = w_eff.clone().requires_grad_(True)
            y, sums = ops.linear(x, w, bias, sparsity=s, want_stats=True)
            go = torch.randn(1024, 1600, generator=torch.Generator().manual_seed(98)).to(cuda).bfloat16()
            gx, gw = torch.autograd.grad(y, [x, w], go)
            outs.append((y, gx, gw, sums))
        (y0, gx0, gw0, s0), (y1, gx1, gw1, s1) = outs
        _check(y1, y0.float(), tol=1e-2)
        _check(gx1, gx0.float(), tol=1e-2)
        _check(gw1 * support, gw0 * support, tol=1e-4)
        assert torch.allclose(s0.sum(0), s1.sum(0), rtol=1e-3, atol=1.0)
    finally:
        gemm_sm100.uninstall()


@pytest.mark.parametrize("defer", [False, True])
def test_graph_conv_op_matches_einsum_formulation(cuda, defer):
    """ops.graph_conv (weight-build kernel + block-sparse GEMM + gradient-fold kernel) against the reference's own
    formulation -- conv 64 -> K*64 followed by einsum 'nkctv,kvw->nctw' (stgcn_layers.py:58-67) -- in fp32 autograd on
    the same bf16-rounded input: output, input gradient, and the gradients of conv weight, conv bias, edge importance."""
    import numpy as np
    from pose2room_b200 import gemm_sm100, ops
    from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency
    gemm_sm100.install()
    try:
        g = torch.Generator().manual_seed(21)
        A = torch.tensor(np.array(spatial_adjacency(layout_for_joints(25), max_hop=5)), dtype=torch.float32).to(cuda)
        K, V, C, N, T = A.shape[0], 25, 64, 2, 256
        sp = gemm_sm100.BlockSparsity((A.abs().sum(0) > 0).t().cpu().numpy())
        conv_w = (torch.randn(K * C, C, 1, 1, generator=g) / 8).to(cuda).requires_grad_(True)
        conv_b = (torch.randn(K * C, generator=g) / 4).to(cuda).requires_grad_(True)
        imp = (1 + 0.3 * torch.randn(K, V, V, generator=g)).to(cuda).requires_grad_(True)
        x = torch.randn(N * T, V * C, generator=g).to(cuda).bfloat16().requires_grad_(True)
        go = torch.randn(N * T, V * C, generator=g).to(cuda).bfloat16()

        def run_ours():
            y, sums = ops.graph_conv(x, conv_w, conv_b, A * imp, sp)
            y.backward(go)
            return y, sums
        if defer:
            with ops.overlap_weight_grads():
                y, sums = run_ours()
        else:
            y, sums = run_ours()
        torch.cuda.synchronize()
        got = [y.detach().float(), x.grad.float(), conv_w.grad.clone(), conv_b.grad.clone(), imp.grad.clone()]
        # reference formulation, fp32
        cw, cb, im = [t.detach().clone().requires_grad_(True) for t in (conv_w, conv_b, imp)]
        xr = x.detach().float().requires_grad_(True)
        xin = xr.reshape(N, T, V, C).permute(0, 3, 1, 2)                            # (N, C, T, V)
        h = torch.nn.functional.conv2d(xin, cw, cb).view(N, K, C, T, V)
        ref = torch.einsum("nkctv,kvw->nctw", h, A * im).permute(0, 2, 3, 1).reshape(N * T, V * C)
        ref.backward(go.float())
        want = [ref.detach(), xr.grad, cw.grad, cb.grad, im.grad]
        for name, a, b, tol in zip(["y", "dx", "dW", "db", "dImportance"], got, want, [1e-2, 1e-2, 1e-2, 1e-2, 1e-2]):
            err, scale = (a - b).abs().max().item(), b.abs().max().item()
            assert err <= tol * scale + 1e-4, (name, err, scale)
        yd = y.detach().double().reshape(N * T * V, C)
        assert torch.allclose(sums.sum(0)[0], yd.sum(0), rtol=1e-6, atol=1e-2)
    finally:
        gemm_sm100.uninstall()


@pytest.mark.parametrize("bn", [128, 256])
def test_gemm_cta_pair_kernel(cuda, bn):
    """The cta_group::2 kernel (one MMA across two SMs): exact on integer operands, ragged M / N, bias + ReLU, block-
    sparse k-lists and fused statistics."""
    import numpy as np
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(30 + bn)
    # the last two shapes give every pair several tiles: the ring and the double-buffered accumulator wrap around
    for (M, N, K) in [(256, bn, 64), (512, 640, 256), (300, 320, 192), (1024, 1600, 1600), (8192, 640, 128),
                      (16384 + 100, 1600, 192)]:
        a = torch.randint(-3, 4, (M, K), generator=g).float()
        b = torch.randint(-3, 4, (N, K), generator=g).float()
        bias = torch.randint(-5, 6, (N,), generator=g).float()
        c = gemm_sm100.gemm_pair(a.to(cuda).bfloat16(), b.to(cuda).bfloat16(), bias=bias.to(cuda), relu=True, block_n=bn)
        ref = torch.relu(a @ b.t() + bias)
        assert torch.equal(c.float().cpu(), ref.bfloat16().float()), (M, N, K)
    # block-sparse reduction + statistics on random data
    M, NB, KB = 768, 10, 10
    nz = _random_block_pattern(NB, KB, g)
    sp = gemm_sm100.BlockSparsity(nz)
    w = torch.randn(NB * 64, KB * 64, generator=g) / 20
    w *= torch.from_numpy(np.kron(nz, np.ones((64, 64), dtype=np.float32)))
    x = torch.randn(M, KB * 64, generator=g)
    xb, wb = x.to(cuda).bfloat16(), w.to(cuda).bfloat16()
    stats = torch.zeros(4, 2, 64, dtype=torch.float64, device=cuda)
    c = gemm_sm100.gemm_pair(xb, wb, block_n=bn, kb_list=sp.kb_list(bn, False, cuda), stats=stats)
    _check(c, xb.float() @ wb.float().t(), tol=1e-2)
    cd = c.double().reshape(M, NB, 64)
    assert torch.allclose(stats.sum(0)[0], cd.sum((0, 1)), rtol=1e-5, atol=1e-2)
    assert torch.allclose(stats.sum(0)[1], (cd * cd).sum((0, 1)), rtol=1e-5, atol=1e-2)


def test_gemm_cta_pair_weight_gradient(cuda):
    """dW = dz^T . x on CTA pairs (MN-major operands, 256 x 256 tiles, split-K by bulk-tensor reduce-add): exact on
    integer operands for the full tile set, and equal to the masked product for a tile list."""
    import numpy as np
    from pose2room_b200 import gemm_sm100
    g = torch.Generator().manual_seed(40)
    for (R, N1, N2, splits) in [(512, 256, 256, 1), (2048, 640, 320, 3), (4096, 1600, 1600, 0), (1000, 384, 200, 2)]:
        dz = torch.randint(-2, 3, (R, N1), generator=g).float()
        x = torch.randint(-2, 3, (R, N2), generator=g).float()
        dw = gemm_sm100.gemm_pair_dw(dz.to(cuda).bfloat16(), x.to(cuda).bfloat16(), splits=splits)
        assert torch.equal(dw.cpu(), dz.t() @ x), (R, N1, N2, splits)
    nz = _random_block_pattern(25, 25, g, density=0.3)
    sp = gemm_sm100.BlockSparsity(nz)
    R = 8192
    dz = torch.randint(-2, 3, (R, 1600), generator=g).float()
    x = torch.randint(-2, 3, (R, 1600), generator=g).float()
    tl = sp.tile_list(256, 256, cuda)
    dw = gemm_sm100.gemm_pair_dw(dz.to(cuda).bfloat16(), x.to(cuda).bfloat16(), tl)
    keep = torch.from_numpy(np.kron(sp.tile_mask(256, 256, "cpu").numpy(), np.ones((256, 256), dtype=np.float32)))[:1600, :1600]
    assert torch.equal(dw.cpu(), (dz.t() @ x) * keep)
    assert (keep.numpy() >= np.kron(nz, np.ones((64, 64), dtype=np.float32))).all()


def test_graph_conv_residual_gradient_fold(cuda):
    """The residual branch's gradient folded into the input-gradient GEMM (reduce-add stores of the CTA-pair kernel: dx is
    ADDED onto the residual gradient in place) against autograd's own sum of the two gradients: M = 8192 rows (the pair
    kernel's path) and a ragged M; also C += A.B^T of gemm_pair directly against a float64 product."""
    import numpy as np
    from pose2room_b200 import gemm_sm100, ops
    from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency
    gemm_sm100.install()
    try:
        g = torch.Generator().manual_seed(5)
        # (1) the GEMM itself
        for m in (8192, 4096 + 300):
            a = torch.randn(m, 1600, generator=g).to(cuda).bfloat16()
            b = (torch.randn(1600, 1600, generator=g) / 40).to(cuda).bfloat16()
            c0 = torch.randn(m, 1600, generator=g).to(cuda).bfloat16()
            c = c0.clone()
            out = gemm_sm100.gemm_pair(a, b, accumulate_into=c)
            assert out.data_ptr() == c.data_ptr()
            want = c0.double() + a.double() @ b.double().t()
            err = (c.double() - want).abs().max().item()
            assert err <= 2e-2 * want.abs().max().item(), err          # two bf16 roundings (the product tile, then the sum)
        # (2) through the operator: x feeds the graph convolution AND a residual branch
        A = torch.tensor(np.array(spatial_adjacency(layout_for_joints(25), max_hop=5)), dtype=torch.float32).to(cuda)
        K, V, C, M = A.shape[0], 25, 64, 8192
        sp = gemm_sm100.BlockSparsity((A.abs().sum(0) > 0).t().cpu().numpy())
        conv_w = (torch.randn(K * C, C, 1, 1, generator=g) / 8).to(cuda).requires_grad_(True)
        conv_b = (torch.randn(K * C, generator=g) / 4).to(cuda).requires_grad_(True)
        x0 = torch.randn(M, V * C, generator=g).to(cuda).bfloat16()
        go, gr = (torch.randn(M, V * C, generator=g).to(cuda).bfloat16() for _ in range(2))
        grads = []
        for fold in (True, False):
            ops._FUSED_RESADD = fold
            x = x0.clone().requires_grad_(True)
            y, _, x_res = ops.graph_conv(x, conv_w, conv_b, A, sp, residual_alias=True)
            torch.autograd.backward([y, x_res * 1.0], [go, gr.clone()])
            grads.append(x.grad.double())
            conv_w.grad = conv_b.grad = None
        ops._FUSED_RESADD = True
        scale = grads[1].abs().max().item()
        assert (grads[0] - grads[1]).abs().max().item() <= 1e-2 * scale
        assert (grads[0] - grads[1]).abs().mean().item() <= 1e-3 * scale
    finally:
        ops._FUSED_RESADD = True
        gemm_sm100.uninstall()
