"""GPU: the dense channel-last operators (GEMM, BatchNorm fwd/bwd, temporal conv, grouping, max-pool) against
plain PyTorch fp32 references of the same ops, forward and backward."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-4, atol=1e-5):
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a.float() - b.float()).abs().max().item()
    assert torch.allclose(a.float(), b.float(), rtol=rtol, atol=atol), err


@pytest.mark.parametrize("M,K,N,bias,relu", [(1000, 64, 64, True, False), (777, 3, 64, False, False),
                                             (300, 256, 259, True, False), (513, 1600, 1600, True, False),
                                             (4096, 256, 256, True, True), (65, 128, 100, True, False)])
def test_linear_fwd_bwd_fp32(cuda, M, K, N, bias, relu):
    from pose2room_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).to(cuda).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda).requires_grad_(True)
    b = torch.randn(N, generator=g).to(cuda).requires_grad_(True) if bias else None
    y = ops.linear(x, w, b, relu=relu)
    ref = F.linear(x.double(), w.double(), b.double() if bias else None)
    ref = F.relu(ref) if relu else ref
    _close(y, ref, rtol=1e-5, atol=1e-5)
    go = torch.randn(M, N, generator=g).to(cuda)
    grads = torch.autograd.grad(y, [x, w] + ([b] if bias else []), go)
    rgrads = torch.autograd.grad(ref, [x, w] + ([b] if bias else []), go.double())
    for a, r in zip(grads, rgrads):
        _close(a, r, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("relu,res", [(True, False), (False, False), (True, True)])
@pytest.mark.parametrize("M,C", [(5000, 64), (333, 256), (4096, 128)])
def test_batchnorm_act_train_fwd_bwd(cuda, M, C, relu, res):
    from pose2room_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(cuda).requires_grad_(True)
    r = torch.randn(M, C, generator=g).to(cuda).requires_grad_(True) if res else None
    bn = nn.BatchNorm1d(C).to(cuda)
    with torch.no_grad():
        bn.weight.normal_(1, 0.2)
        bn.bias.normal_(0, 0.2)
    bn_ref = nn.BatchNorm1d(C).to(cuda).double()
    bn_ref.load_state_dict(bn.state_dict())
    y = ops.batchnorm_act(x, bn, relu=relu, residual=r)
    xr = x.detach().double().requires_grad_(True)
    rr = r.detach().double().requires_grad_(True) if res else None
    ref = bn_ref(xr)
    if res:
        ref = ref + rr
    if relu:
        ref = F.relu(ref)
    _close(y, ref, rtol=1e-4, atol=1e-5)
    _close(bn.running_mean, bn_ref.running_mean, rtol=1e-5, atol=1e-6)
    _close(bn.running_var, bn_ref.running_var, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1
    go = torch.randn(M, C, generator=g).to(cuda)
    y.backward(go)
    ref.backward(go.double())
    # fp32 partial sums + atomics in arbitrary order: compare against the fp64 reference relative to each tensor's scale
    def close_scaled(a, b, tol):
        scale = b.abs().max().item() + 1e-12
        assert (a.double() - b).abs().max().item() <= tol * scale, ((a.double() - b).abs().max().item(), scale)
    close_scaled(x.grad, xr.grad, 2e-4)
    close_scaled(bn.weight.grad, bn_ref.weight.grad, 2e-4)
    close_scaled(bn.bias.grad, bn_ref.bias.grad, 2e-4)
    if res:
        _close(r.grad, rr.grad, rtol=1e-5, atol=1e-6)


def test_batchnorm_eval_mode(cuda):
    from pose2room_b200 import ops
    bn = nn.BatchNorm1d(64).to(cuda)
    with torch.no_grad():
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2)
        bn.weight.normal_(1, 0.2)
        bn.bias.normal_()
    bn.eval()
    x = torch.randn(1000, 64, device=cuda, requires_grad=True)
    y = ops.batchnorm_act(x, bn, relu=True)
    xr = x.detach().clone().requires_grad_(True)
    ref = F.relu(bn(xr))
    _close(y, ref, rtol=1e-5, atol=1e-6)
    go = torch.randn_like(y)
    y.backward(go)
    ref.backward(go)
    _close(x.grad, xr.grad, rtol=1e-5, atol=1e-6)


def test_temporal_conv_fwd_bwd(cuda):
    from pose2room_b200 import ops
    B, T, V, C = 2, 37, 25, 64
    conv = nn.Conv2d(C, C, (3, 1), (1, 1), (1, 0)).to(cuda)
    x = torch.randn(B, T, V, C, device=cuda, requires_grad=True)
    y = ops.temporal_conv(x, conv.weight, conv.bias).reshape(B, T, V, C)
    xr = x.detach().clone().requires_grad_(True)
    ref = conv(xr.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    _close(y, ref, rtol=1e-4, atol=1e-5)
    go = torch.randn_like(y)
    gw, gb, gx = torch.autograd.grad(y, [conv.weight, conv.bias, x], go)
    rw, rb, rx = torch.autograd.grad(ref, [conv.weight, conv.bias, xr], go)
    _close(gx, rx, rtol=1e-4, atol=1e-5)
    _close(gw, rw, rtol=1e-3, atol=1e-4)
    _close(gb, rb, rtol=1e-3, atol=1e-4)


def test_group_rows_and_maxpool(cuda):
    from pose2room_b200 import ops
    B, N, C, P, S = 3, 100, 256, 16, 16
    feats = torch.randn(B, N, C, device=cuda, requires_grad=True)
    idx = torch.randint(0, N, (B, P, S), device=cuda, dtype=torch.int32)
    out = ops.group_rows(feats, idx)
    fr = feats.detach().clone().requires_grad_(True)
    ref = torch.gather(fr[:, None].expand(B, P, N, C), 2, idx.long()[..., None].expand(B, P, S, C))
    assert torch.equal(out, ref)
    pooled = ops.maxpool_rows(out.reshape(B * P, S, C))
    rp = ref.reshape(B * P, S, C).max(dim=1).values
    assert torch.equal(pooled, rp)
    go = torch.randn_like(pooled)
    pooled.backward(go)
    rp.backward(go)
    _close(feats.grad, fr.grad, rtol=1e-5, atol=1e-5)


def test_bf16_storage_variants(cuda):
    """Same kernels with bf16 activations (fp32 arithmetic): agree with fp32 up to bf16 rounding."""
    from pose2room_b200 import ops
    x = torch.randn(2048, 64, device=cuda)
    w = torch.randn(64, 64, device=cuda) / 8
    y32 = ops.linear(x, w)
    y16 = ops.linear(x.bfloat16(), w)
    assert y16.dtype == torch.bfloat16
    _close(y16, y32, rtol=2e-2, atol=2e-2)
    bn = nn.BatchNorm1d(64).to(cuda)
    z16 = ops.batchnorm_act(y16, bn, relu=True)
    bn2 = nn.BatchNorm1d(64).to(cuda)
    z32 = ops.batchnorm_act(y32, bn2, relu=True)
    _close(z16, z32, rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("M", [4096, 5000 * 3 + 17, 70000])
@pytest.mark.parametrize("relu,res", [(True, False), (False, False), (True, True)])
def test_streaming_batchnorm_bf16_matches_fp64(cuda, M, relu, res):
    """[M, 64] bf16 operands take the bulk-TMA streaming kernels (stream_bn.cu); forward, running statistics and all
    gradients against an fp64 BatchNorm on the same bf16-rounded inputs, including a ragged last tile."""
    from pose2room_b200 import ops
    C = 64
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, C, generator=g) * 2 + 0.5).to(cuda).bfloat16().requires_grad_(True)
    r = torch.randn(M, C, generator=g).to(cuda).bfloat16().requires_grad_(True) if res else None
    bn = nn.BatchNorm1d(C).to(cuda)
    with torch.no_grad():
        bn.weight.normal_(1, 0.2)
        bn.bias.normal_(0, 0.2)
    bn_ref = nn.BatchNorm1d(C).to(cuda).double()
    bn_ref.load_state_dict(bn.state_dict())
    y = ops.batchnorm_act(x, bn, relu=relu, residual=r)
    assert y.dtype == torch.bfloat16
    xr = x.detach().double().requires_grad_(True)
    rr = r.detach().double().requires_grad_(True) if res else None
    ref = bn_ref(xr)
    if res:
        ref = ref + rr
    if relu:
        ref = F.relu(ref)
    _close(y, ref, rtol=1e-2, atol=1e-2)                                  # bf16 storage of y
    _close(bn.running_mean, bn_ref.running_mean, rtol=1e-5, atol=1e-6)    # statistics are exact sums of the bf16 inputs
    _close(bn.running_var, bn_ref.running_var, rtol=1e-5, atol=1e-6)
    go = torch.randn(M, C, generator=g).to(cuda).bfloat16()
    # the ReLU mask of the kernels comes from the bf16 y (with residual) or the fp32 pre-activation (without): mask the
    # reference gradient with the kernels' own y so that isolated sign flips at |pre-activation| ~ 0 do not count
    y.backward(go)
    if relu:
        mask = (y.detach().double() > 0).double()
        pre = bn_ref(xr) + (rr if res else 0)
        (pre * mask).backward(go.double())
    else:
        ref.backward(go.double())

    def close_scaled(a, b, tol):
        scale = b.abs().max().item() + 1e-12
        assert (a.double() - b).abs().max().item() <= tol * scale, ((a.double() - b).abs().max().item(), scale)
    close_scaled(x.grad, xr.grad, 1e-2)
    close_scaled(bn.weight.grad, bn_ref.weight.grad, 2e-3)
    close_scaled(bn.bias.grad, bn_ref.bias.grad, 2e-3)
    if res:
        close_scaled(r.grad, rr.grad, 1e-2)


def test_streaming_kernels_equal_generic_kernels(cuda):
    """The same C-ABI calls with P2R_STREAM_BN toggled (subprocess: the switch is read once) give identical column sums
    (up to fp32 partial-sum order) and bit-identical elementwise outputs."""
    import json, os, subprocess, sys
    code = r'''
import json, torch
from pose2room_b200 import _lib
dev = torch.device("cuda:0"); M, C = 50000 + 37, 64
g = torch.Generator().manual_seed(5)
x = torch.randn(M, C, generator=g).to(dev).bfloat16(); dy = torch.randn(M, C, generator=g).to(dev).bfloat16()
yy = torch.randn(M, C, generator=g).to(dev).bfloat16()
st = torch.rand(4, C, generator=g).to(dev) + 0.5
s = torch.zeros(6, C, dtype=torch.float64, device=dev)
out = torch.zeros(5, M, C, dtype=torch.bfloat16, device=dev)
cs = torch.cuda.current_stream().cuda_stream
_lib.call("p2r_col_stats", x.data_ptr(), 1, M, C, s[0].data_ptr(), s[1].data_ptr(), cs)
_lib.call("p2r_col_bwd_stats", dy.data_ptr(), x.data_ptr(), None, 1, M, C, st[0].data_ptr(), st[1].data_ptr(), 2, s[2].data_ptr(), s[3].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), cs)
_lib.call("p2r_col_bwd_stats", dy.data_ptr(), x.data_ptr(), yy.data_ptr(), 1, M, C, st[0].data_ptr(), st[1].data_ptr(), 1, s[4].data_ptr(), s[5].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), cs)
_lib.call("p2r_affine_act", x.data_ptr(), 1, M, C, st[2].data_ptr(), st[3].data_ptr(), None, 1, out[0].data_ptr(), None, cs)
_lib.call("p2r_affine_act", x.data_ptr(), 1, M, C, st[2].data_ptr(), st[3].data_ptr(), yy.data_ptr(), 1, out[1].data_ptr(), None, cs)
_lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), None, 1, M, C, st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), s[2].data_ptr(), s[3].data_ptr(), 2, out[2].data_ptr(), None, st[3].data_ptr(), None, 0, cs)
_lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), yy.data_ptr(), 1, M, C, st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), s[4].data_ptr(), s[5].data_ptr(), 1, out[3].data_ptr(), out[4].data_ptr(), st[3].data_ptr(), None, 0, cs)
torch.cuda.synchronize()
print(json.dumps({"s": s.cpu().tolist(), "o": [float(out[i].float().sum()) for i in range(5)],
                  "h": [int(out[i].view(torch.int16).long().sum()) for i in range(5)]}))
'''
    res = []
    for flag in ("0", "1"):
        env = dict(os.environ, P2R_STREAM_BN=flag)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    a, b = res
    sa, sb = torch.tensor(a["s"]), torch.tensor(b["s"])
    assert torch.allclose(sa, sb, rtol=1e-5, atol=1e-2), (sa - sb).abs().max()
    assert a["h"][0] == b["h"][0] and a["h"][1] == b["h"][1] and a["h"][4] == b["h"][4]   # bit-identical outputs
    # dx depends on the sums (tiny differences in the last fp32 bit of the coefficients): compare as values
    assert abs(a["o"][2] - b["o"][2]) <= 1e-3 * (abs(b["o"][2]) + 1) and abs(a["o"][3] - b["o"][3]) <= 1e-3 * (abs(b["o"][3]) + 1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed_sum_fwd_bwd(cuda, dtype):
    """x = sk + mean_k(pos) broadcast over the joints of a frame (stgcn.py:121,129); bf16 with C = 64 takes the
    vectorised kernel.  Integer-valued data keeps bf16 exact."""
    from pose2room_b200 import ops
    g = torch.Generator().manual_seed(9)
    F_, J, K, C = 777, 25, 20, 64
    sk = torch.randint(-3, 4, (F_, J, C), generator=g).float().to(cuda).to(dtype).requires_grad_(True)
    pos = (torch.randint(-2, 3, (F_, K, C), generator=g).float() * K).to(cuda).to(dtype).requires_grad_(True)   # mean stays integral
    x = ops.embed_sum(sk, pos)
    ref = sk.detach().float() + pos.detach().float().mean(1, keepdim=True)
    assert torch.equal(x.float(), ref)
    go = (torch.randint(-2, 3, (F_, J, C), generator=g).float() * 4).to(cuda).to(dtype)
    gsk, gpos = torch.autograd.grad(x, [sk, pos], go)
    assert torch.equal(gsk.float(), go.float())
    want = (go.float().sum(1, keepdim=True) / K).expand(F_, K, C)
    assert torch.allclose(gpos.float(), want, rtol=1e-2 if dtype == torch.bfloat16 else 1e-6, atol=1e-6)


@pytest.mark.parametrize("P", [25, 1])
def test_bn_backward_fused_periodic_column_sums(cuda, P):
    """p2r_bn_bwd_apply with colsum / period: the per-(row % period, channel) sums of the dx it writes equal a separate
    column sum over the stored bf16 dx (period 25: the graph convolution's bias gradient over the 25 joints, shared-memory
    table; period 1: the temporal conv's bias gradient, per-thread register sums)."""
    from pose2room_b200 import _lib
    if _lib.query("p2r_stream_bn_supported", 1, 25 * 2048, 64) != 1:
        pytest.skip("streaming kernels disabled")
    M, C = 25 * 2048 + 25 * 3, 64
    g = torch.Generator().manual_seed(77)
    x = torch.randn(M, C, generator=g).to(cuda).bfloat16()
    dy = torch.randn(M, C, generator=g).to(cuda).bfloat16()
    st = (torch.rand(4, C, generator=g) + 0.5).to(cuda)
    s = torch.zeros(2, C, dtype=torch.float64, device=cuda)
    cs_stream = torch.cuda.current_stream().cuda_stream
    _lib.call("p2r_col_bwd_stats", dy.data_ptr(), x.data_ptr(), None, 1, M, C, st[0].data_ptr(), st[1].data_ptr(), 2,
              s[0].data_ptr(), s[1].data_ptr(), st[2].data_ptr(), st[3].data_ptr(), cs_stream)
    dx = torch.empty_like(x)
    colsum = torch.zeros(P, C, dtype=torch.float64, device=cuda)
    _lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), None, 1, M, C, st[0].data_ptr(), st[1].data_ptr(),
              st[2].data_ptr(), s[0].data_ptr(), s[1].data_ptr(), 2, dx.data_ptr(), None, st[3].data_ptr(),
              colsum.data_ptr(), P, cs_stream)
    want = dx.double().reshape(M // P, P, C).sum(0)
    assert torch.allclose(colsum, want, rtol=1e-4, atol=1e-2), (colsum - want).abs().max()


@pytest.mark.parametrize("M,K", [(70017, 3), (4096, 3), (1000, 4), (70020, 3), (8196, 4)])      # the last three take the bulk-TMA ring
@pytest.mark.parametrize("training", [True, False])
def test_embed_first_layer_without_stored_preactivation(cuda, M, K, training):
    """ops.embed_l1 (conv K -> 64 + BatchNorm + ReLU on float32 coordinates; statistics from the moments of the input, z
    recomputed in the backward, dW accumulated in the pass that computes dz: csrc/embed_ops.cu) against an fp64 torch
    run of the same three modules: output, running statistics, gradients of the conv weight and of gamma / beta."""
    from pose2room_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    x = (torch.randn(M, K, generator=g) * torch.tensor([0.3, 1.0, 2.0, 0.5][:K]) + torch.tensor([0.1, -0.9, 0.4, 2.0][:K])).to(cuda)
    w = (torch.randn(64, K, generator=g) / K ** 0.5).to(cuda).requires_grad_(True)
    bn = nn.BatchNorm1d(64).to(cuda)
    with torch.no_grad():
        bn.weight.normal_(1, 0.2)
        bn.bias.normal_(0, 0.2)
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 1.5)
    bn_ref = nn.BatchNorm1d(64).to(cuda).double()
    bn_ref.load_state_dict(bn.state_dict())
    bn.train(training)
    bn_ref.train(training)
    assert ops.embed_l1_ok(x, 64, K)
    y = ops.embed_l1(x, w, bn)
    assert y.dtype == torch.bfloat16 and y.shape == (M, 64)
    wr = w.detach().double().requires_grad_(True)
    ref = F.relu(bn_ref(x.double() @ wr.t()))
    assert float((y.double() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())            # bf16 output
    assert float((y.double() - ref).norm() / ref.norm()) <= 3e-3
    _close(bn.running_mean, bn_ref.running_mean, rtol=1e-4, atol=1e-5)
    _close(bn.running_var, bn_ref.running_var, rtol=1e-4, atol=1e-5)
    assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)
    go = torch.randn(M, 64, generator=g).to(cuda).bfloat16()
    gw, gg, gb = torch.autograd.grad(y, [w, bn.weight, bn.bias], go)
    # the reference masks with ITS ReLU; entries within bf16 rounding of zero may differ: compare norms
    rw, rg, rb = torch.autograd.grad(ref, [wr, bn_ref.weight, bn_ref.bias], go.double())
    for name, a, r in (("dW", gw, rw), ("dgamma", gg, rg), ("dbeta", gb, rb)):
        assert float((a.double() - r).norm() / (r.norm() + 1e-12)) <= 2e-3, (name, float((a.double() - r).norm() / r.norm()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,N,C,P", [(3, 50, 1600, 16), (2, 33, 25, 40), (1, 8, 64, 1), (2, 3, 40, 150)])
def test_select_rows_and_its_destination_major_adjoint(cuda, dtype, B, N, C, P):
    """ops.select_rows (seed-frame pick before conv_joint, stgcn.py:136-139) against torch.gather, forward and backward,
    with frames picked twice (P > N in the second case) and never: the adjoint stores every row and sums duplicates."""
    from pose2room_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + C)
    feats = torch.randn(B, N, C, generator=g).to(cuda).to(dtype).requires_grad_(True)
    idx = torch.randint(0, N, (B, P), generator=g).to(cuda)
    out = ops.select_rows(feats, idx)
    want = torch.gather(feats, 1, idx[:, :, None].expand(B, P, C))
    assert torch.equal(out, want)
    go = torch.randn(B, P, C, generator=g).to(cuda).to(dtype)
    (got,) = torch.autograd.grad(out, feats, go)
    ref = torch.zeros(B, N, C, dtype=torch.float64, device=cuda)
    ref.scatter_add_(1, idx[:, :, None].expand(B, P, C), go.double())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert got.dtype == dtype and torch.allclose(got.double(), ref, rtol=tol, atol=tol)
    untouched = torch.ones(B, N, dtype=torch.bool, device=cuda)
    untouched.scatter_(1, idx, False)
    assert float(got[untouched].abs().max()) == 0.0 if bool(untouched.any()) else True


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C", [(8192, 259), (8192, 100), (1000, 24), (37, 7), (300, 1000)])
@pytest.mark.parametrize("relu", [False, True])
def test_bias_gradient_column_sums_of_any_width(cuda, dtype, M, C, relu):
    """ops._col_sum on widths the thread-per-column kernels do not take (the 259-column vote head, the 100 mixture
    weights, the 24 box parameters: proposal_module.py / vote_center.py heads) against a float64 sum."""
    from pose2room_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    dy = torch.randn(M, C, generator=g).to(cuda).to(dtype)
    y = torch.randn(M, C, generator=g).to(cuda).to(dtype) if relu else None
    got = ops._col_sum(dy, y, relu)
    want = (dy.double() * (y > 0) if relu else dy.double()).sum(0)
    assert got.dtype == torch.float32 and got.shape == (C,)
    assert torch.allclose(got.double(), want, rtol=1e-4, atol=1e-3 * M ** 0.5)


def test_bn_backward_writes_float_sums_in_the_apply_launch(cuda):
    """p2r_bn_bwd_apply_ex: sums32 = float(sums64), dx identical to p2r_bn_bwd_apply (streaming and fallback widths)."""
    from pose2room_b200 import _lib
    for M, C, dt in ((4096, 64, torch.bfloat16), (513, 32, torch.float32)):
        g = torch.Generator().manual_seed(C)
        x = torch.randn(M, C, generator=g).to(cuda).to(dt)
        dy = torch.randn(M, C, generator=g).to(cuda).to(dt)
        st = torch.rand(4, C, generator=g).to(cuda) + 0.5
        sums = (torch.randn(2, C, generator=g) * 100).double().to(cuda)
        code = {torch.float32: 0, torch.bfloat16: 1}[dt]
        cs = torch.cuda.current_stream().cuda_stream
        dx0, dx1 = torch.empty_like(x), torch.empty_like(x)
        s32 = torch.full((2, C), float("nan"), device=cuda)
        _lib.call("p2r_bn_bwd_apply", dy.data_ptr(), x.data_ptr(), None, code, M, C, st[0].data_ptr(), st[1].data_ptr(),
                  st[2].data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), 2, dx0.data_ptr(), None, st[3].data_ptr(), None, 0, cs)
        _lib.call("p2r_bn_bwd_apply_ex", dy.data_ptr(), x.data_ptr(), None, code, M, C, st[0].data_ptr(), st[1].data_ptr(),
                  st[2].data_ptr(), sums[0].data_ptr(), sums[1].data_ptr(), 2, dx1.data_ptr(), None, st[3].data_ptr(), None, 0,
                  sums.data_ptr(), s32.data_ptr(), cs)
        assert torch.equal(dx0, dx1)
        assert torch.equal(s32, sums.float())
