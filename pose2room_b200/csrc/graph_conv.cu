// graph_conv.cu -- weight plumbing of the fused graph convolution of st_gcn_block.
//
// The reference computes  y = einsum('nkctv,kvw->nctw', conv1x1_{64 -> K*64}(x), A)   (stgcn_layers.py:58-67, A = the
// adjacency stack times the learned edge importance, stgcn.py:133-134).  Here it is ONE GEMM per block against
//      W_eff[(w,co), (v,ci)] = sum_k A[k,v,w] * W[k*Co+co, ci],      b_eff[(w,co)] = sum_k b[k*Co+co] * sum_v A[k,v,w]
// (see gemm_sm100.cu).  These two kernels build W_eff / its transpose / b_eff in bf16 straight from the conv parameters
// (forward) and fold the GEMM's weight gradient dW_eff back onto the conv weight, the conv bias and A (backward), so the
// step contains no einsum / bmm / transpose / dtype-conversion glue kernels for them.
#include "p2r_common.cuh"

#define GC_C 64   // channels per joint block (Co = Ci = 64 on the hot path)

// grid (V, V): block (w = blockIdx.x, v = blockIdx.y), 256 threads; thread t owns row r = t / 4, columns 16 (t % 4) .. +15
__global__ void __launch_bounds__(256)
gcn_build_weight_kernel(const float* __restrict__ conv_w, const float* __restrict__ conv_b, const float* __restrict__ A,
                        int K, int V, __nv_bfloat16* __restrict__ w_eff, __nv_bfloat16* __restrict__ w_eff_t,
                        float* __restrict__ b_eff) {
  __shared__ float tile[GC_C][GC_C + 1];
  const int w = blockIdx.x, v = blockIdx.y, t = threadIdx.x;
  const int r = t >> 2, c0 = (t & 3) * 16;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float a = __ldg(A + ((size_t)k * V + v) * V + w);
    if (a != 0.f) {                                         // (block-uniform)
      const float* src = conv_w + ((size_t)k * GC_C + r) * GC_C + c0;     // W[k*Co + co = r][ci = c0..]
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src + i));
        acc[i] = fmaf(a, q.x, acc[i]);
        acc[i + 1] = fmaf(a, q.y, acc[i + 1]);
        acc[i + 2] = fmaf(a, q.z, acc[i + 2]);
        acc[i + 3] = fmaf(a, q.w, acc[i + 3]);
      }
    }
  }
  // W_eff block rows w*64 + co, columns v*64 + ci
  const size_t ld = (size_t)V * GC_C;
  {
    __nv_bfloat16* dst = w_eff + ((size_t)w * GC_C + r) * ld + (size_t)v * GC_C + c0;
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  // transpose through shared memory: W_eff^T block rows v*64 + ci, columns w*64 + co
#pragma unroll
  for (int i = 0; i < 16; ++i) tile[r][c0 + i] = acc[i];
  __syncthreads();
  {
    __nv_bfloat16* dst = w_eff_t + ((size_t)v * GC_C + r) * ld + (size_t)w * GC_C + c0;   // here r = ci, c0.. = co
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(tile[c0 + 2 * i][r], tile[c0 + 2 * i + 1][r]);
      pk[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
  if (v == 0 && t < GC_C && b_eff != nullptr) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
      float cs = 0.f;
      for (int u = 0; u < V; ++u) cs += __ldg(A + ((size_t)k * V + u) * V + w);
      s = fmaf(conv_b ? __ldg(conv_b + k * GC_C + t) : 0.f, cs, s);
    }
    b_eff[w * GC_C + t] = s;
  }
}

// conv_w [K*64, 64] fp32 (row k*64 + co), conv_b [K*64] fp32 or NULL, A [K, V, V] fp32 ->
// w_eff [V*64, V*64] bf16, w_eff_t (its transpose) bf16, b_eff [V*64] fp32.
extern "C" int p2r_gcn_build_weight(const float* conv_w, const float* conv_b, const float* A, int K, int V, int Co, int Ci,
                                    void* w_eff, void* w_eff_t, float* b_eff, void* stream) {
  P2R_CHECK_ARG(K > 0 && V > 0 && Co == GC_C && Ci == GC_C, "p2r_gcn_build_weight (built for 64 -> 64 channel blocks)");
  P2R_LAUNCH(gcn_build_weight_kernel, dim3(V, V), 256, 0, (cudaStream_t)stream, conv_w, conv_b, A, K, V,
             (__nv_bfloat16*)w_eff, (__nv_bfloat16*)w_eff_t, b_eff);
  P2R_RETURN_LAUNCH("p2r_gcn_build_weight");
}

// Backward of the construction above, two launches, no atomics (deterministic):
//   gcn_dA_kernel, grid (V, V), block (w, v):
//     dA[k,v,w] = <W_k, dW_eff block (w,v)> + sum_co conv_b[k,co] * db_eff[w,co]   where A[k,v,w] != 0, else 0
//   gcn_dw_kernel, grid (K, 64), block (k, output row co), 8 warps that split the joints v between them (warp i takes
//   v = i, i + 8, ...: eight short chains of dependent loads instead of one long one -- round 1's version, one warp per
//   row walking all V*V pairs, took 68 us for 10 MB), combined through shared memory in warp order:
//     d_conv_w[k*64+co, ci] = sum over (v,w) with A[k,v,w] != 0 of A[k,v,w] * dW_eff[(w,co),(v,ci)]
//     d_conv_b[k*64+co]     = sum_w db_eff[w,co] * sum_v A[k,v,w]
// Entries of dA where A == 0 are written as 0: A = adjacency * importance, so the chain rule multiplies them by the
// adjacency's zero anyway, and the structurally-zero blocks of dW_eff are never computed (tile mask of the dW GEMM).
__global__ void __launch_bounds__(256)
gcn_dA_kernel(const float* __restrict__ dw_eff, const float* __restrict__ db_eff, const float* __restrict__ conv_w,
              const float* __restrict__ conv_b, const float* __restrict__ A, int K, int V, float* __restrict__ dA) {
  __shared__ float red[8];
  const int w = blockIdx.x, v = blockIdx.y, t = threadIdx.x;
  const int r = t >> 2, c0 = (t & 3) * 16;
  const int warp = t >> 5, lane = t & 31;
  const size_t ld = (size_t)V * GC_C;
  float blk[16];
  bool loaded = false;
  for (int k = 0; k < K; ++k) {
    const float a = __ldg(A + ((size_t)k * V + v) * V + w);
    if (a == 0.f) {                                          // (block-uniform)
      if (t == 0) dA[((size_t)k * V + v) * V + w] = 0.f;
      continue;
    }
    if (!loaded) {
      const float* src = dw_eff + ((size_t)w * GC_C + r) * ld + (size_t)v * GC_C + c0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src + i));
        blk[i] = q.x; blk[i + 1] = q.y; blk[i + 2] = q.z; blk[i + 3] = q.w;
      }
      loaded = true;
    }
    const float* wk = conv_w + ((size_t)k * GC_C + r) * GC_C + c0;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) dot = fmaf(__ldg(wk + i), blk[i], dot);
    if (t < GC_C && conv_b != nullptr && db_eff != nullptr)
      dot = fmaf(__ldg(conv_b + k * GC_C + t), __ldg(db_eff + w * GC_C + t), dot);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    __syncthreads();                                         // red[] free (previous k consumed)
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    if (t == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i];
      dA[((size_t)k * V + v) * V + w] = s;
    }
  }
}

__global__ void __launch_bounds__(256)
gcn_dw_kernel(const float* __restrict__ dw_eff, const float* __restrict__ db_eff, const float* __restrict__ A, int K,
              int V, float* __restrict__ d_conv_w, float* __restrict__ d_conv_b) {
  P2R_DYN_SMEM(float, sA);                                   // A[k] : V x V, then red[8][64]
  float* red = sA + V * V;
  const int k = blockIdx.x, co = blockIdx.y, t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  for (int i = t; i < V * V; i += 256) sA[i] = __ldg(A + (size_t)k * V * V + i);
  __syncthreads();
  const int ci = lane * 2;                                   // two adjacent columns per lane: 256-byte rows per warp
  const size_t ld = (size_t)V * GC_C;
  float acc0 = 0.f, acc1 = 0.f;
  for (int v = warp; v < V; v += 8) {
    const float* base = dw_eff + (size_t)co * ld + (size_t)v * GC_C + ci;
#pragma unroll 4
    for (int w = 0; w < V; ++w) {
      const float a = sA[v * V + w];
      if (a != 0.f) {                                        // (warp-uniform)
        const float2 q = __ldg(reinterpret_cast<const float2*>(base + (size_t)w * GC_C * ld));
        acc0 = fmaf(a, q.x, acc0);
        acc1 = fmaf(a, q.y, acc1);
      }
    }
  }
  red[warp * GC_C + ci] = acc0;
  red[warp * GC_C + ci + 1] = acc1;
  __syncthreads();
  if (t < GC_C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i * GC_C + t];      // fixed order: deterministic
    d_conv_w[((size_t)k * GC_C + co) * GC_C + t] = s;
  }
  if (d_conv_b != nullptr && db_eff != nullptr && warp == 1) {
    float s = 0.f;
    for (int w = lane; w < V; w += 32) {
      float cs = 0.f;
      for (int v = 0; v < V; ++v) cs += sA[v * V + w];
      s = fmaf(__ldg(db_eff + w * GC_C + co), cs, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) d_conv_b[k * GC_C + co] = s;
  }
}

extern "C" int p2r_gcn_reduce_weight_grad(const float* dw_eff, const float* db_eff, const float* conv_w,
                                          const float* conv_b, const float* A, int K, int V, int Co, int Ci,
                                          float* d_conv_w, float* d_conv_b, float* dA, void* stream) {
  P2R_CHECK_ARG(K > 0 && V > 0 && Co == GC_C && Ci == GC_C, "p2r_gcn_reduce_weight_grad (64 -> 64 channel blocks)");
  P2R_CHECK_ARG(((size_t)V * V + 8 * GC_C) * sizeof(float) <= 48 * 1024, "p2r_gcn_reduce_weight_grad (adjacency too large for shared memory)");
  P2R_LAUNCH(gcn_dA_kernel, dim3(V, V), 256, 0, (cudaStream_t)stream, dw_eff, db_eff, conv_w, conv_b, A, K, V, dA);
  P2R_LAUNCH(gcn_dw_kernel, dim3(K, GC_C), 256, ((size_t)V * V + 8 * GC_C) * sizeof(float), (cudaStream_t)stream, dw_eff,
             db_eff, A, K, V, d_conv_w, d_conv_b);
  P2R_RETURN_LAUNCH("p2r_gcn_reduce_weight_grad");
}
