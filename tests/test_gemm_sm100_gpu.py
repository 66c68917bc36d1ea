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
        for (M, K, N) in [(2048, 1600, 1600), (8192, 64, 64), (4096, 192, 64), (1024, 256, 256)]:
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
