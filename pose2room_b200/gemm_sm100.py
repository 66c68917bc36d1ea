"""Throughput-mode GEMM backend: routes the dense layers' bf16 GEMMs to the tcgen05/TMA kernel of
csrc/gemm_sm100.cu (C ABI `p2r_gemm_bf16`).  `install()` hooks it into pose2room_b200.ops.linear so that
forward (x.W^T), input gradient (dy.W) and weight gradient (dy^T.x) of every layer with bf16 activations run on
the tensor cores; layers the kernel cannot take (K or N not a multiple of 8, tiny K) stay on the SIMT kernel."""
import os
import weakref

import numpy as np
import torch

from . import _lib, ops

# tiling of the block-sparse graph-convolution GEMMs (forward / input gradient) and number of partial-statistics copies
GCN_BLOCK_N = int(os.environ.get("P2R_GCN_BLOCK_N", "128"))
STAT_COPIES = int(os.environ.get("P2R_STAT_COPIES", "16"))
USE_SPARSITY = os.environ.get("P2R_GCN_SPARSE", "1") != "0"
# graph-conv forward / input-gradient GEMMs on CTA pairs (persistent cta_group::2 kernel, 256 x PAIR_BLOCK_N tiles)
USE_PAIR = os.environ.get("P2R_GCN_PAIR", "1") != "0"
PAIR_BLOCK_N = int(os.environ.get("P2R_GCN_PAIR_BLOCK_N", "256"))
# Weight gradient of the graph convolution on CTA pairs (gemm2_dw_kernel: 256 x 256 tiles, 124 us in the step against ~250
# for the 128 x 128-tile kernel, which is bound by L2 -> SM operand traffic).  Round 1 and the first half of round 2 kept it
# off: with the backward's two streams balanced as they were then, a persistent kernel that owns the shared memory / TMEM
# of its SMs cost more on the BatchNorm chain than it saved (9.44-9.55 ms/step vs 9.32).  After the rest of the backward got
# shorter the weight-gradient stream became the longer one, and the faster kernel wins: 7.44 vs 7.55 ms/step.
DW_TILE = (256, 256) if os.environ.get("P2R_GCN_PAIR_DW", "1") != "0" else (128, 128)
USE_PAIR_DW = os.environ.get("P2R_GCN_PAIR_DW", "1") != "0"
USE_FUSED_STATS = os.environ.get("P2R_FUSED_STATS", "1") != "0"


class BlockSparsity:
    """64x64 block pattern of a weight W[N,K] (nn.Linear layout): nz[nb, kb] is False where the block is structurally
    zero.  For the graph convolution W_eff[(w,co),(v,ci)] = sum_k W_k[co,ci] A_k[v,w] (ref: stgcn_layers.py:58-67 with
    A from Graph.get_adjacency, :163-208) block (w,v) is zero whenever joints v and w are more than max_hop apart in
    the skeleton -- 48 % of the blocks of the 25-joint layout.  Produces the k-block lists / tile masks that
    p2r_gemm_bf16_ex consumes, cached per device."""

    def __init__(self, nz):
        self.nz = np.asarray(nz, dtype=bool)
        self._cache = {}

    @property
    def density(self):
        return float(self.nz.mean())

    def kb_list(self, block_n, transposed, device):
        """int32 [tiles_n, 1 + kblocks]: row i = (count, k-block indices...) of n-tile i.  transposed=False: forward,
        n runs over W's rows; True: input gradient dx = dz.W, n runs over W's columns and k over its rows."""
        key = ("kb", block_n, transposed, str(device))
        if key not in self._cache:
            nz = self.nz.T if transposed else self.nz          # [n blocks, k blocks]
            n_cols = nz.shape[0] * 64
            tiles_n = -(-n_cols // block_n)
            tab = np.zeros((tiles_n, 1 + nz.shape[1]), dtype=np.int32)
            for i in range(tiles_n):
                b0, b1 = (i * block_n) // 64, (min(n_cols, (i + 1) * block_n) - 1) // 64
                ks = np.nonzero(nz[b0:b1 + 1].any(0))[0]
                tab[i, 0] = len(ks)
                tab[i, 1:1 + len(ks)] = ks
            self._cache[key] = torch.from_numpy(tab).to(device)
        return self._cache[key]

    def tile_list(self, block_m, block_n, device):
        """int32 [T, 2]: (row tile, column tile) of the structurally non-zero block_m x block_n tiles of dW[N,K]."""
        key = ("tiles", block_m, block_n, str(device))
        if key not in self._cache:
            mask = self.tile_mask(block_m, block_n, "cpu").numpy()
            self._cache[key] = torch.from_numpy(np.argwhere(mask > 0).astype(np.int32)).contiguous().to(device)
        return self._cache[key]

    def tile_mask(self, block_m, block_n, device):
        """uint8 [tiles_m, tiles_n] for the weight gradient dW[N,K] tiled block_m x block_n."""
        key = ("mask", block_m, block_n, str(device))
        if key not in self._cache:
            n, k = self.nz.shape[0] * 64, self.nz.shape[1] * 64
            tm, tn = -(-n // block_m), -(-k // block_n)
            mask = np.zeros((tm, tn), dtype=np.uint8)
            for i in range(tm):
                for j in range(tn):
                    r0, r1 = (i * block_m) // 64, (min(n, (i + 1) * block_m) - 1) // 64
                    c0, c1 = (j * block_n) // 64, (min(k, (j + 1) * block_n) - 1) // 64
                    mask[i, j] = self.nz[r0:r1 + 1, c0:c1 + 1].any()
            self._cache[key] = torch.from_numpy(mask).to(device)
        return self._cache[key]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def available():
    if not torch.cuda.is_available():
        return False
    major, _ = torch.cuda.get_device_capability()
    return major == 10 and hasattr(_lib.load(), "p2r_gemm_bf16")


def gemm(a, b, a_mn=False, b_mn=False, bias=None, relu=False, out_dtype=torch.bfloat16, splits=1, block_n=0,
         kb_list=None, tile_mask=None, stats=None, zero_skipped=True):
    """C[M,N] = op(a) @ op(b)^T.  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn); bf16, contiguous rows.
    kb_list / tile_mask / stats: the extras of p2r_gemm_bf16_ex (block-sparse reduction, skipped output tiles --
    those stay zero, or unwritten with zero_skipped=False for a consumer that never reads them --, fused per-channel
    output statistics [copies, 2, 64] float64, zero-filled here by the caller)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
    assert a.stride(1) == 1 and b.stride(1) == 1
    k, m = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
    kb, n = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
    assert k == kb, (a.shape, b.shape, a_mn, b_mn)
    if splits > 1:
        c = torch.zeros(m, n, dtype=torch.float32, device=a.device)
    elif tile_mask is not None and zero_skipped:
        c = torch.zeros(m, n, dtype=out_dtype, device=a.device)
    else:
        c = torch.empty(m, n, dtype=out_dtype, device=a.device)
    if bias is not None:
        bias = bias.float().contiguous()
    with torch.cuda.device(a.device):
        if kb_list is None and tile_mask is None and stats is None:
            _lib.call("p2r_gemm_bf16", m, n, k, a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn),
                      c.data_ptr(), c.stride(0), 1 if c.dtype == torch.bfloat16 else 0,
                      bias.data_ptr() if bias is not None else None, int(relu), int(splits), int(block_n), _stream())
        else:
            if kb_list is not None:
                assert kb_list.dtype == torch.int32 and kb_list.is_contiguous() and kb_list.shape[0] == -(-n // block_n)
            if tile_mask is not None:
                assert tile_mask.dtype == torch.uint8 and tile_mask.is_contiguous()
                assert tuple(tile_mask.shape) == (-(-m // 128), -(-n // block_n))
            if stats is not None:
                assert stats.dtype == torch.float64 and stats.is_contiguous() and tuple(stats.shape[1:]) == (2, 64)
            _lib.call("p2r_gemm_bf16_ex", m, n, k, a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0),
                      int(b_mn), c.data_ptr(), c.stride(0), 1 if c.dtype == torch.bfloat16 else 0,
                      bias.data_ptr() if bias is not None else None, int(relu), int(splits), int(block_n),
                      kb_list.data_ptr() if kb_list is not None else None,
                      int(kb_list.shape[1]) if kb_list is not None else 0,
                      tile_mask.data_ptr() if tile_mask is not None else None,
                      stats.data_ptr() if stats is not None else None, int(stats.shape[0]) if stats is not None else 1,
                      _stream())
    return c


def gemm_pair(a, b, bias=None, relu=False, block_n=256, kb_list=None, stats=None, _debug_flags=0, accumulate_into=None):
    """C[M,N] bf16 = a[M,K] @ b[N,K]^T on CTA pairs (p2r_gemm_bf16_pair): K-major bf16 operands, 256 x block_n tiles.
    accumulate_into: a contiguous bf16 [M,N] tensor the product is ADDED to in place (reduce-add stores) and returned."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
    assert a.stride(1) == 1 and b.stride(1) == 1 and a.shape[1] == b.shape[1]
    m, k = a.shape
    n = b.shape[0]
    if accumulate_into is not None:
        c = accumulate_into
        assert c.dtype == torch.bfloat16 and c.is_contiguous() and tuple(c.shape) == (m, n) and bias is None and not relu
        _debug_flags |= 2
    else:
        c = torch.empty(m, n, dtype=torch.bfloat16, device=a.device)
    if bias is not None:
        bias = bias.float().contiguous()
    if kb_list is not None:
        assert kb_list.dtype == torch.int32 and kb_list.is_contiguous() and kb_list.shape[0] == -(-n // block_n)
    if stats is not None:
        assert stats.dtype == torch.float64 and stats.is_contiguous() and tuple(stats.shape[1:]) == (2, 64)
    # (per-kernel timing for bench.py's roofline leg: the events bracket this launch alone)
    with torch.cuda.device(a.device), ops._Timed("pair_fwd" if stats is not None else "pair", m, n, k):
        _lib.call("p2r_gemm_bf16_pair", m, n, k, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), c.data_ptr(),
                  c.stride(0), bias.data_ptr() if bias is not None else None, int(relu) | int(_debug_flags), int(block_n),
                  kb_list.data_ptr() if kb_list is not None else None,
                  int(kb_list.shape[1]) if kb_list is not None else 0,
                  stats.data_ptr() if stats is not None else None, int(stats.shape[0]) if stats is not None else 1,
                  _stream())
    return c


def gemm_pair_dw(dz, x, tile_list=None, splits=0):
    """dW[N1,N2] fp32 = dz[R,N1]^T @ x[R,N2] on CTA pairs (p2r_gemm_bf16_pair_dw): 256 x 256 tiles from `tile_list`
    (default: all), split-K combined by bulk-tensor reduce-add stores."""
    assert dz.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and dz.shape[0] == x.shape[0]
    assert dz.stride(1) == 1 and x.stride(1) == 1
    r, n1 = dz.shape
    n2 = x.shape[1]
    if tile_list is None:
        tm, tn = -(-n1 // 256), -(-n2 // 256)
        tile_list = torch.tensor([[i, j] for i in range(tm) for j in range(tn)], dtype=torch.int32, device=dz.device)
    assert tile_list.dtype == torch.int32 and tile_list.is_contiguous() and tile_list.shape[1] == 2
    t = tile_list.shape[0]
    if splits <= 0:      # ~5 waves of work units over the 74 CTA pairs
        splits = max(1, min(round(370 / t), r // 1024))
    dw = torch.zeros(n1, n2, dtype=torch.float32, device=dz.device)
    with torch.cuda.device(dz.device):
        _lib.call("p2r_gemm_bf16_pair_dw", r, n1, n2, dz.data_ptr(), dz.stride(0), x.data_ptr(), x.stride(0),
                  dw.data_ptr(), dw.stride(0), tile_list.data_ptr(), t, int(splits), _stream())
    return dw


_TCONV_PREBUILT = {}
ops._STEP_END_HOOKS.append(_TCONV_PREBUILT.clear)


def _tconv_layouts(weight, bias):
    """The two bf16 layouts of a (Co, Ci, KT, 1) temporal-conv weight the implicit GEMMs read -- forward [Co, KT*Ci]
    (column dt*Ci + ci), input gradient [KT*Co, Ci] (row dt*Co + co) -- and the float32 bias."""
    co, ci, kt, _ = weight.shape
    wb = ops.bf16_weight(weight)[:, :, :, 0]
    w2 = wb.permute(0, 2, 1).reshape(co, kt * ci).contiguous()
    wt = wb.permute(2, 0, 1).reshape(kt * co, ci).contiguous()
    return w2, wt, (bias.float().contiguous() if bias is not None else None)


def tconv_prebuild(convs):
    """Lay the temporal-conv weights out NOW, on the current stream (a forked branch at the top of the forward: parameters
    only), for the next temporal_conv() call with the same weight tensor; unused entries are dropped when the step ends."""
    _TCONV_PREBUILT.clear()
    out = []
    with torch.no_grad():
        for weight, bias in convs:
            built = _tconv_layouts(weight.detach(), bias.detach() if bias is not None else None)
            _TCONV_PREBUILT[weight.data_ptr()] = built
            out.append(built)
    if ops.DEFER["on"]:
        ops.DEFER["keep"].append(out)
    return out


class _TemporalConvTC(torch.autograd.Function):
    """(KT x 1) temporal conv of st_gcn_block.tcn as implicit tensor-core GEMMs (forward, d input, d weight):
    three row-shifted TMA views of the SAME activation tensor instead of an unfold buffer."""

    @staticmethod
    def forward(ctx, x, weight, bias, targets=None, want_stats=False):
        ctx.targets = targets
        b, t, v, ci = x.shape
        co, _, kt, _ = weight.shape
        x = x if x.is_contiguous() else x.contiguous()
        w2, wt, bias_f = _TCONV_PREBUILT.pop(weight.data_ptr(), None) or _tconv_layouts(weight, bias)
        ctx.wt = wt
        y = torch.empty(b * t * v, co, dtype=torch.bfloat16, device=x.device)
        sums = None
        if want_stats and USE_FUSED_STATS and co == 64:
            sums = ops.zeros_ws((STAT_COPIES, 2, 64), torch.float64, x.device)
        with torch.cuda.device(x.device):
            _lib.call("p2r_tconv_bf16", 0, x.data_ptr(), w2.data_ptr(), None, y.data_ptr(), b, t * v, ci, co, kt, v,
                      bias_f.data_ptr() if bias_f is not None else None, 1,
                      sums.data_ptr() if sums is not None else None, STAT_COPIES, _stream())
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        if want_stats:
            if sums is None:
                sums = torch.empty(0, dtype=torch.float64, device=x.device)
            ctx.mark_non_differentiable(sums)
            ctx.set_materialize_grads(False)
            return y, sums
        return y

    @staticmethod
    def backward(ctx, dy, _dsums=None):
        if dy is None:
            return (None,) * 5
        x, weight = ctx.saved_tensors
        b, t, v, ci = x.shape
        co, _, kt, _ = weight.shape
        dy = dy if dy.is_contiguous() else dy.contiguous()
        # [1, Co] sums of dy left by the backward of the BatchNorm that follows this conv (ops.batchnorm_act(...,
        # colsum_period=1)): the bias gradient without another pass over dy
        fused_cs = ops._COLSUM.pop(dy.data_ptr(), None)
        if fused_cs is not None and fused_cs.numel() != co:
            fused_cs = None
        dx = dw = db = None
        with torch.cuda.device(x.device):
            if ctx.needs_input_grad[0]:
                wt = ctx.wt
                dx = torch.empty(b, t, v, ci, dtype=torch.bfloat16, device=x.device)
                _lib.call("p2r_tconv_bf16", 1, dy.data_ptr(), wt.data_ptr(), None, dx.data_ptr(), b, t * v, ci, co, kt, v,
                          None, 1, None, 1, _stream())
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
                                  kt, v, None, max(1, min(148, m // 4096)), None, 1, _stream())
                        gw = dw2.reshape(co, kt, ci).permute(0, 2, 1).unsqueeze(-1)
                    if need_b:
                        gb = ops._to_float_at_join(fused_cs.reshape(-1)) if fused_cs is not None else ops._col_sum(dy, at_join=True)
                    return [gw, gb]
                ops._defer(weight_grads, [tw if need_w else None, tb if need_b else None], (dy, x))
                return dx, None, None, None, None
            if ctx.needs_input_grad[1]:
                m = b * t * v
                splits = max(1, min(148, m // 4096))
                dw2 = torch.zeros(co, kt * ci, dtype=torch.float32, device=x.device)
                _lib.call("p2r_tconv_bf16", 2, x.data_ptr(), None, dy.data_ptr(), dw2.data_ptr(), b, t * v, ci, co, kt, v,
                          None, splits, None, 1, _stream())
                dw = dw2.reshape(co, kt, ci).permute(0, 2, 1).unsqueeze(-1)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = ops._col_sum(dy)
        return dx, dw, db, None, None


def _pad8(n):
    return (n + 7) // 8 * 8


def _pad_rows(t, rows):
    """[n, k] -> [rows, k] with zero rows appended (odd output widths: 259-channel vote head, 100 mixture weights).
    A registered bf16 weight shadow already lives in a zero-padded buffer (ops.register_weight_shadows): no copy."""
    parent = ops.padded_shadow(t, rows)
    if parent is not None:
        return parent
    out = torch.zeros(rows, t.shape[1], dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


_PAD_MEMO = {"key": None, "value": None, "src": None}


def _pad_cols(t, cols):
    """[m, n] -> [m, cols] with zero columns appended.  The input gradient and the weight gradient of an odd-width layer
    both pad the same dz, one right after the other: the last result is kept and reused while its source is alive and
    unchanged (same storage, same version counter)."""
    key = (t.data_ptr(), tuple(t.shape), cols, t._version, t.dtype)
    if _PAD_MEMO["key"] == key and _PAD_MEMO["src"]() is t:
        return _PAD_MEMO["value"]
    out = torch.zeros(t.shape[0], cols, dtype=t.dtype, device=t.device)
    out[:, :t.shape[1]] = t
    _PAD_MEMO["key"], _PAD_MEMO["value"], _PAD_MEMO["src"] = key, out, weakref.ref(t)
    if ops.DEFER["on"]:
        ops.DEFER["keep"].append(out)      # the deferred weight gradient reads it on the side stream: alive until the join
    return out


class _Backend:
    """Interface expected by ops._Linear (see ops._TC_GEMM)."""
    # the CTA-pair weight-gradient kernel on the weight-gradient stream like every other dW (measured at the end of round 2:
    # 7.44 ms/step there, 7.78 in line with the backward chain, 7.55 with the 128 x 128-tile kernel)
    DW_INLINE = USE_PAIR_DW and os.environ.get("P2R_GCN_DW_INLINE", "0") != "0"

    @staticmethod
    def supports(m, n, k):
        # TMA needs 16-byte row pitches for every operand in every role (x:[M,K], w:[N,K], dy:[M,N]); an output width
        # that is not a multiple of 8 is zero-padded to the next one (the extra columns are sliced off again)
        return k % 8 == 0 and k >= 32 and m >= 128 and n >= 8

    @staticmethod
    def linear_fwd(x, weight, bias, relu):
        n = weight.shape[0]
        if n % 8:
            n8 = _pad8(n)
            w8 = _pad_rows(weight.to(torch.bfloat16), n8)
            b8 = torch.cat([bias.float(), bias.new_zeros(n8 - n, dtype=torch.float32)]) if bias is not None else None
            return gemm(x, w8, False, False, b8, relu, out_dtype=torch.bfloat16)[:, :n].contiguous()
        return gemm(x, weight.to(torch.bfloat16), False, False, bias, relu, out_dtype=torch.bfloat16)

    @staticmethod
    def linear_fwd_ex(x, weight, bias, relu, sparsity, want_stats):
        """forward with the block-sparse reduction and / or the fused BatchNorm statistics of the output."""
        n = weight.shape[0]
        sums = None
        if want_stats and USE_FUSED_STATS and n % 64 == 0:
            sums = ops.zeros_ws((STAT_COPIES, 2, 64), torch.float64, x.device)
        sp = sparsity if USE_SPARSITY else None
        if sp is None and sums is None:
            return _Backend.linear_fwd(x, weight, bias, relu), None
        if USE_PAIR and x.shape[0] >= 4096 and n >= 512 and n % 64 == 0:
            kbl = sp.kb_list(PAIR_BLOCK_N, False, x.device) if sp is not None else None
            y = gemm_pair(x, weight.to(torch.bfloat16), bias, relu, block_n=PAIR_BLOCK_N, kb_list=kbl, stats=sums)
            return y, sums
        if n <= 64:
            bn = 64
        else:
            bn = GCN_BLOCK_N if (n % 160 == 0 or GCN_BLOCK_N != 160) else 128
        kbl = sp.kb_list(bn, False, x.device) if sp is not None else None
        y = gemm(x, weight.to(torch.bfloat16), False, False, bias, relu, out_dtype=torch.bfloat16, block_n=bn,
                 kb_list=kbl, stats=sums)
        return y, sums

    @staticmethod
    def linear_dx(dz, weight, sparsity=None):
        # dx[M,K] = dz[M,N] . W[N,K]:  A = dz (K-major over n), B = W viewed as [K_red = N, N_out = K] -> MN-major
        w = weight.to(torch.bfloat16)
        if w.shape[0] % 8:      # odd layer width: zero-pad the reduction dimension of both operands
            n8 = _pad8(w.shape[0])
            dz, w = _pad_cols(dz, n8), _pad_rows(w, n8)
        n_out = w.shape[1]
        sp = sparsity if USE_SPARSITY else None
        # (the CTA-pair kernel on the one wide dense layer, conv_joint's 256 -> 1600 input gradient, measured 7.04 ms per
        # step against 7.00 with the 160-wide tiles: four k-blocks do not amortise its pipeline)
        if (n_out % 160 == 0 and n_out >= 640) or sp is not None:
            # graph-conv sized layers: a 5 MB transpose of W_eff buys the zero-waste 128x160 K-major tiling
            bn = GCN_BLOCK_N if sp is not None else 160
            kbl = sp.kb_list(bn, True, dz.device) if sp is not None else None
            return gemm(dz, w.t().contiguous(), False, False, out_dtype=torch.bfloat16, block_n=bn, kb_list=kbl)
        return gemm(dz, w, False, True, out_dtype=torch.bfloat16)

    @staticmethod
    def linear_dx_pretransposed(dz, w_t, sparsity=None, accumulate_into=None):
        """dx = dz . W with W^T [K, N] already materialised in bf16 (the graph-conv weight builder writes both).
        accumulate_into: add dx onto this tensor in place (CTA-pair kernel only) and return it."""
        sp = sparsity if USE_SPARSITY else None
        if USE_PAIR and dz.shape[0] >= 4096 and w_t.shape[0] >= 512:
            kbl = sp.kb_list(PAIR_BLOCK_N, True, dz.device) if sp is not None else None
            return gemm_pair(dz, w_t, block_n=PAIR_BLOCK_N, kb_list=kbl, accumulate_into=accumulate_into)
        if accumulate_into is not None:
            return accumulate_into.add_(_Backend.linear_dx_pretransposed(dz, w_t, sparsity))
        bn = GCN_BLOCK_N if (w_t.shape[0] % 160 == 0 or GCN_BLOCK_N != 160) else 128
        kbl = sp.kb_list(bn, True, dz.device) if sp is not None else None
        return gemm(dz, w_t, False, False, out_dtype=torch.bfloat16, block_n=bn, kb_list=kbl)

    @staticmethod
    def linear_dw(dz, x, sparsity=None, zero_skipped=True):
        # dW[N,K] = dz^T[N,M] . x[M,K]: both operands MN-major (reduction over the row index m)
        if dz.shape[1] % 8:     # odd layer width: pad the columns of dz, drop the extra rows of dW
            n_true = dz.shape[1]
            return _Backend.linear_dw(_pad_cols(dz, _pad8(n_true)), x, sparsity)[:n_true]
        m = dz.shape[0]
        n, k = dz.shape[1], x.shape[1]
        tiles = ((n + 127) // 128) * ((k + 127) // 128)
        if USE_PAIR_DW and n >= 1024 and k >= 1024 and m >= 8192:   # graph conv: CTA pairs, 256 x 256 tiles, reduce-add split-K
            tl = sparsity.tile_list(256, 256, dz.device) if (sparsity is not None and USE_SPARSITY) else None
            return gemm_pair_dw(dz, x, tl)
        if tiles >= 148:  # graph-conv: 13 x 13 = 169 tiles of 128x128 fill the chip without split-K / atomics
            mask = sparsity.tile_mask(128, 128, dz.device) if (sparsity is not None and USE_SPARSITY) else None
            return gemm(dz, x, True, True, out_dtype=torch.float32, splits=1, block_n=128, tile_mask=mask,
                        zero_skipped=zero_skipped)
        # split-K over about one CTA per SM (measured, tools/dw_splits_bench.py: 819200 x 64 x 64 56 us at 148 splits, 62 at
        # 200, 69 at 296; 16384 x 256 x 256 24 us at 16, 33 at 4; below 8192 rows the zero-fill + atomics cost more than they buy)
        # (below 8192 rows one CTA per tile: split-K for the 4096-row box heads -- ~40 us each as a serial chain on the
        # weight-gradient stream -- was measured at 7.17-7.21 ms per step against 7.00 without)
        splits = 1 if m < 8192 else max(1, min((148 + tiles - 1) // tiles, m // 1024, 148 if tiles == 1 else 32))
        return gemm(dz, x, True, True, out_dtype=torch.float32, splits=splits)

    @staticmethod
    def supports_tconv(shape, co):
        b, t, v, ci = shape
        return ci == 64 and co == 64 and (t * v) % 128 == 0

    @staticmethod
    def temporal_conv(x, weight, bias, want_stats=False):
        if ops.DEFER["on"] and torch.is_grad_enabled() and x.requires_grad:
            return _TemporalConvTC.apply(x, weight.detach(), bias.detach() if bias is not None else None, (weight, bias),
                                         want_stats)
        return _TemporalConvTC.apply(x, weight, bias, None, want_stats)


def install():
    if not available():
        raise RuntimeError("pose2room_b200.gemm_sm100: needs an sm_100 device and libp2r_b200.so")
    ops._TC_GEMM["fn"] = _Backend
    return _Backend


def uninstall():
    ops._TC_GEMM["fn"] = None
