"""Throughput-mode GEMM backend: routes the dense layers' bf16 GEMMs to the tcgen05/TMA kernel of
csrc/gemm_sm100.cu (C ABI `p2r_gemm_bf16`).  `install()` hooks it into pose2room_b200.ops.linear so that
forward (x.W^T), input gradient (dy.W) and weight gradient (dy^T.x) of every layer with bf16 activations run on
the tensor cores; layers the kernel cannot take (K or N not a multiple of 8, tiny K) stay on the SIMT kernel."""
import torch

from . import _lib, ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


def available():
    if not torch.cuda.is_available():
        return False
    major, _ = torch.cuda.get_device_capability()
    return major == 10 and hasattr(_lib.load(), "p2r_gemm_bf16")


def gemm(a, b, a_mn=False, b_mn=False, bias=None, relu=False, out_dtype=torch.bfloat16, splits=1, block_n=0):
    """C[M,N] = op(a) @ op(b)^T.  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn); bf16, contiguous rows."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
    assert a.stride(1) == 1 and b.stride(1) == 1
    k, m = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
    kb, n = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
    assert k == kb, (a.shape, b.shape, a_mn, b_mn)
    if splits > 1:
        c = torch.zeros(m, n, dtype=torch.float32, device=a.device)
    else:
        c = torch.empty(m, n, dtype=out_dtype, device=a.device)
    if bias is not None:
        bias = bias.float().contiguous()
    with torch.cuda.device(a.device):
        _lib.call("p2r_gemm_bf16", m, n, k, a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn),
                  c.data_ptr(), c.stride(0), 1 if c.dtype == torch.bfloat16 else 0,
                  bias.data_ptr() if bias is not None else None, int(relu), int(splits), int(block_n), _stream())
    return c


class _TemporalConvTC(torch.autograd.Function):
    """(KT x 1) temporal conv of st_gcn_block.tcn as implicit tensor-core GEMMs (forward, d input, d weight):
    three row-shifted TMA views of the SAME activation tensor instead of an unfold buffer."""

    @staticmethod
    def forward(ctx, x, weight, bias, targets=None):
        ctx.targets = targets
        b, t, v, ci = x.shape
        co, _, kt, _ = weight.shape
        x = x if x.is_contiguous() else x.contiguous()
        w2 = weight[:, :, :, 0].permute(0, 2, 1).reshape(co, kt * ci).to(torch.bfloat16).contiguous()
        y = torch.empty(b * t * v, co, dtype=torch.bfloat16, device=x.device)
        bias_f = bias.float().contiguous() if bias is not None else None
        with torch.cuda.device(x.device):
            _lib.call("p2r_tconv_bf16", 0, x.data_ptr(), w2.data_ptr(), None, y.data_ptr(), b, t * v, ci, co, kt, v,
                      bias_f.data_ptr() if bias_f is not None else None, 1, _stream())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        b, t, v, ci = x.shape
        co, _, kt, _ = weight.shape
        dy = dy if dy.is_contiguous() else dy.contiguous()
        dx = dw = db = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                wt = weight[:, :, :, 0].permute(2, 0, 1).reshape(kt * co, ci).to(torch.bfloat16).contiguous()
                dx = torch.empty(b, t, v, ci, dtype=torch.bfloat16, device=x.device)
                _lib.call("p2r_tconv_bf16", 1, dy.data_ptr(), wt.data_ptr(), None, dx.data_ptr(), b, t * v, ci, co, kt, v,
                          None, 1, _stream())
            if ctx.targets is not None:
                tw, tb = ctx.targets
                need_w = tw is not None and tw.requires_grad
                need_b = tb is not None and tb.requires_grad

                def weight_grads():
                    gw = gb = None
                    if need_w:
                        m = b * t * v
                        dw2 = torch.zeros(co, kt * ci, dtype=torch.float32, device=x.device)
                        _lib.call("p2r_tconv_bf16", 2, x.data_ptr(), None, dy.data_ptr(), dw2.data_ptr(), b, t * v, ci, co,
                                  kt, v, None, max(1, min(128, m // 8192)), _stream())
                        gw = dw2.reshape(co, kt, ci).permute(0, 2, 1).unsqueeze(-1)
                    if need_b:
                        gb = ops._col_sum(dy)
                    return [gw, gb]
                ops._defer(weight_grads, [tw if need_w else None, tb if need_b else None], (dy, x))
                return dx, None, None, None
            if ctx.needs_input_grad[1]:
                m = b * t * v
                splits = max(1, min(128, m // 8192))
                dw2 = torch.zeros(co, kt * ci, dtype=torch.float32, device=x.device)
                _lib.call("p2r_tconv_bf16", 2, x.data_ptr(), None, dy.data_ptr(), dw2.data_ptr(), b, t * v, ci, co, kt, v,
                          None, splits, _stream())
                dw = dw2.reshape(co, kt, ci).permute(0, 2, 1).unsqueeze(-1)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = ops._col_sum(dy)
        return dx, dw, db, None


class _Backend:
    """Interface expected by ops._Linear (see ops._TC_GEMM)."""

    @staticmethod
    def supports(m, n, k):
        # TMA needs 16-byte row pitches for every operand in every role (x:[M,K], w:[N,K], dy:[M,N])
        return k % 8 == 0 and n % 8 == 0 and k >= 32 and m >= 128

    @staticmethod
    def linear_fwd(x, weight, bias, relu):
        return gemm(x, weight.to(torch.bfloat16), False, False, bias, relu, out_dtype=torch.bfloat16)

    @staticmethod
    def linear_dx(dz, weight):
        # dx[M,K] = dz[M,N] . W[N,K]:  A = dz (K-major over n), B = W viewed as [K_red = N, N_out = K] -> MN-major
        w = weight.to(torch.bfloat16)
        n_out = w.shape[1]
        if n_out % 160 == 0 and n_out >= 640:
            # graph-conv sized layers: a 5 MB transpose of W_eff buys the zero-waste 128x160 K-major tiling
            return gemm(dz, w.t().contiguous(), False, False, out_dtype=torch.bfloat16, block_n=160)
        return gemm(dz, w, False, True, out_dtype=torch.bfloat16)

    @staticmethod
    def linear_dw(dz, x):
        # dW[N,K] = dz^T[N,M] . x[M,K]: both operands MN-major (reduction over the row index m)
        m = dz.shape[0]
        n, k = dz.shape[1], x.shape[1]
        tiles = ((n + 127) // 128) * ((k + 127) // 128)
        if tiles >= 148:  # graph-conv: 13 x 13 = 169 tiles of 128x128 fill the chip without split-K / atomics
            return gemm(dz, x, True, True, out_dtype=torch.float32, splits=1, block_n=128)
        splits = max(1, min((296 + tiles - 1) // tiles, (m + 4095) // 4096))
        return gemm(dz, x, True, True, out_dtype=torch.float32, splits=splits)

    @staticmethod
    def supports_tconv(shape, co):
        b, t, v, ci = shape
        return ci == 64 and co == 64 and (t * v) % 128 == 0

    @staticmethod
    def temporal_conv(x, weight, bias):
        if ops.DEFER["on"] and torch.is_grad_enabled():
            return _TemporalConvTC.apply(x, weight.detach(), bias.detach() if bias is not None else None, (weight, bias))
        return _TemporalConvTC.apply(x, weight, bias)


def install():
    if not available():
        raise RuntimeError("pose2room_b200.gemm_sm100: needs an sm_100 device and libp2r_b200.so")
    ops._TC_GEMM["fn"] = _Backend
    return _Backend


def uninstall():
    ops._TC_GEMM["fn"] = None
