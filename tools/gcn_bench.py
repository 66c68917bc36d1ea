"""Graph-convolution GEMM microbench (diagnostic, not the bench.py contract): dense vs block-sparse k-lists vs fused
statistics, forward / dx / dW, per tile width, next to cuBLAS dense on the same shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root (run as `python tools/<name>.py`)

import json

import numpy as np
import torch

from pose2room_b200 import gemm_sm100
from pose2room_b200.p2rnet.graph import layout_for_joints, spatial_adjacency
from gemm_bench import timeit


def main():
    dev = torch.device("cuda:0")
    M, N = 32768, 1600
    A = np.array(spatial_adjacency(layout_for_joints(25), max_hop=5))
    sp = gemm_sm100.BlockSparsity((np.abs(A).sum(0) > 0).T)
    x = torch.randn(M, N, device=dev).bfloat16()
    w = (torch.randn(N, N, device=dev) / 40).bfloat16()
    dy = torch.randn(M, N, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    rec = {"M": M, "N": N, "K": N, "block_density": sp.density}
    rec["cublas_ms"] = timeit(lambda: x @ w.t())
    for bn in (128, 160, 256):
        rec["dense_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm(x, w, bias=bias, block_n=bn))
        kbl = sp.kb_list(bn, False, dev)
        rec["sparse_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm(x, w, bias=bias, block_n=bn, kb_list=kbl))
        st = torch.zeros(16, 2, 64, dtype=torch.float64, device=dev)
        rec["sparse_stats_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm(x, w, bias=bias, block_n=bn, kb_list=kbl, stats=st))
        rec["dense_stats_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm(x, w, bias=bias, block_n=bn, stats=st))
    for bn in (128, 256):
        kbl = sp.kb_list(bn, False, dev)
        st = torch.zeros(16, 2, 64, dtype=torch.float64, device=dev)
        rec["pair_dense_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm_pair(x, w, bias=bias, block_n=bn))
        rec["pair_sparse_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm_pair(x, w, bias=bias, block_n=bn, kb_list=kbl))
        rec["pair_sparse_stats_bn%d_ms" % bn] = timeit(lambda: gemm_sm100.gemm_pair(x, w, bias=bias, block_n=bn, kb_list=kbl, stats=st))
    rec["dw_dense_ms"] = timeit(lambda: gemm_sm100.gemm(dy, x, True, True, out_dtype=torch.float32, block_n=128))
    mask = sp.tile_mask(128, 128, dev)
    rec["dw_masked_ms"] = timeit(lambda: gemm_sm100.gemm(dy, x, True, True, out_dtype=torch.float32, block_n=128, tile_mask=mask))
    tl = sp.tile_list(256, 256, dev)
    rec["dw_pair_tiles"] = [int(tl.shape[0]), 49]
    for sk in (0, 5, 9, 14):
        rec["dw_pair_s%d_ms" % sk] = timeit(lambda: gemm_sm100.gemm_pair_dw(dy, x, tl, splits=sk))
    rec["dw_pair_dense_ms"] = timeit(lambda: gemm_sm100.gemm_pair_dw(dy, x))
    rec["dw_active_tiles"] = [int(mask.sum()), int(mask.numel())]
    # temporal conv forward with / without statistics
    B, T, V, C = 32, 1024, 25, 64
    from pose2room_b200 import _lib
    h = torch.randn(B, T, V, C, device=dev).bfloat16()
    w2 = (torch.randn(C, 3 * C, device=dev) / 14).bfloat16()
    y = torch.empty(B * T * V, C, device=dev, dtype=torch.bfloat16)
    bt = torch.randn(C, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    st = torch.zeros(16, 2, 64, dtype=torch.float64, device=dev)
    rec["tconv_fwd_ms"] = timeit(lambda: _lib.call("p2r_tconv_bf16", 0, h.data_ptr(), w2.data_ptr(), None, y.data_ptr(), B, T * V, C, C, 3, V, bt.data_ptr(), 1, None, 1, stream))
    rec["tconv_fwd_stats_ms"] = timeit(lambda: _lib.call("p2r_tconv_bf16", 0, h.data_ptr(), w2.data_ptr(), None, y.data_ptr(), B, T * V, C, C, 3, V, bt.data_ptr(), 1, st.data_ptr(), 16, stream))
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
