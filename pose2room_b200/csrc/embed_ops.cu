// embed_ops.cu -- the FIRST layer of the point MLPs of the backbone (pos_embed / sk_feat: Conv1d 3 -> 64, BatchNorm, ReLU;
// ref: models/p2rnet/modules/stgcn.py:45-50 with sub_modules.py:88-113) without ever storing its pre-activation.
//
// The layer maps K <= 4 coordinates to N = 64 channels: z = W x is linear in x, so
//   * the BatchNorm batch statistics of z follow from the first and second MOMENTS of x (K + K*K numbers):
//       mean_c = w_c . E[x],   var_c = w_c^T Cov[x] w_c   -- a 10 MB pass over the coordinates instead of writing,
//       re-reading and normalising a 105 MB [M, 64] tensor;
//   * z is cheaper to recompute from 12 bytes than to load from 128: forward writes a1 = relu(scale z + shift) once
//     (bf16), the backward recomputes z for the ReLU mask and x-hat, and the weight gradient dW = sum_m dz_m x_m^T is
//     accumulated in the pass that computes dz, which is therefore never written either.
// Per MLP and step this replaces (smallk_linear, column statistics, affine pass | BN backward statistics, BN backward
// apply, smallk_dw) = 7 passes over a 105 MB tensor by 1 write + 2 reads.
// Arithmetic: z, the statistics and dz in fp32 / double (round 1 rounded z to bf16 before the BatchNorm: this path is the
// more accurate one); a1 and the incoming gradient are bf16.
#include "p2r_common.cuh"
#include <stdlib.h>

#define EMB_MAXK 4
#define EMB_THREADS 256

__device__ __forceinline__ void emb_load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void emb_store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(p) = v;
}

// ---- moments of the coordinates: s[0..K) = sum x_k, s[K + i*K + j] = sum x_i x_j (double, zero-filled by the caller) ----
__global__ void __launch_bounds__(EMB_THREADS)
coord_moments_kernel(const float* __restrict__ x, long long M, int K, double* __restrict__ s) {
  __shared__ double red[EMB_THREADS / 32][EMB_MAXK + EMB_MAXK * EMB_MAXK];
  double a[EMB_MAXK + EMB_MAXK * EMB_MAXK];
#pragma unroll
  for (int i = 0; i < EMB_MAXK + EMB_MAXK * EMB_MAXK; ++i) a[i] = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += stride) {
    float v[EMB_MAXK] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) v[k] = __ldg(x + m * K + k);
#pragma unroll
    for (int i = 0; i < EMB_MAXK; ++i) {
      a[i] += (double)v[i];
#pragma unroll
      for (int j = 0; j < EMB_MAXK; ++j) a[EMB_MAXK + i * EMB_MAXK + j] += (double)v[i] * (double)v[j];
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < EMB_MAXK + EMB_MAXK * EMB_MAXK; ++i) {
    double t = a[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[warp][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < EMB_MAXK + EMB_MAXK * EMB_MAXK) {
    const int i = threadIdx.x;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < EMB_THREADS / 32; ++w) t += red[w][i];
    // compact layout [K + K*K]
    if (i < EMB_MAXK) { if (i < K) atomicAdd(s + i, t); }
    else {
      const int r = (i - EMB_MAXK) / EMB_MAXK, c = (i - EMB_MAXK) % EMB_MAXK;
      if (r < K && c < K) atomicAdd(s + K + r * K + c, t);
    }
  }
}

extern "C" int p2r_coord_moments(const float* x, long long M, int K, double* s, void* stream) {
  P2R_CHECK_ARG(M >= 0 && K >= 1 && K <= EMB_MAXK, "p2r_coord_moments");
  if (M == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 4, (M + EMB_THREADS - 1) / EMB_THREADS);
  P2R_LAUNCH(coord_moments_kernel, grid, EMB_THREADS, 0, (cudaStream_t)stream, x, M, K, s);
  P2R_RETURN_LAUNCH("p2r_coord_moments");
}

// ---- BatchNorm coefficients of z = W x from the moments (training), running statistics updated like nn.BatchNorm ----
__global__ void embed_l1_finalize_kernel(const double* __restrict__ s, double inv_m, double unbias, int K,
                                         const float* __restrict__ W, int N, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, float momentum,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                                         float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  double mu = 0.0, ez2 = 0.0;
  for (int i = 0; i < K; ++i) {
    const double wi = (double)W[c * K + i];
    mu += wi * s[i] * inv_m;
    for (int j = 0; j < K; ++j) ez2 += wi * (double)W[c * K + j] * s[K + i * K + j] * inv_m;
  }
  double var = ez2 - mu * mu;
  if (var < 0.0) var = 0.0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  rstd[c] = rs;
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  scale[c] = g * rs;
  shift[c] = bt - (float)mu * g * rs;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
  }
}

extern "C" int p2r_embed_l1_finalize(const double* s, long long M, int K, const float* W, int N, const float* gamma,
                                     const float* beta, float eps, float momentum, float* running_mean,
                                     float* running_var, float* mean, float* rstd, float* scale, float* shift,
                                     void* stream) {
  P2R_CHECK_ARG(M > 0 && K >= 1 && K <= EMB_MAXK && N > 0, "p2r_embed_l1_finalize");
  const double unbias = M > 1 ? (double)M / (double)(M - 1) : 1.0;
  P2R_LAUNCH(embed_l1_finalize_kernel, p2r_ceil_div(N, 128), 128, 0, (cudaStream_t)stream, s, 1.0 / (double)M, unbias, K,
             W, N, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift);
  P2R_RETURN_LAUNCH("p2r_embed_l1_finalize");
}

// ---- the three streaming passes.  A thread owns CPT consecutive channels (N / CPT threads per row), so its weights and
// coefficients live in registers for the whole grid-stride loop.  The backward passes take CPT = 4 (8-byte loads, half the
// registers: three CTAs per SM instead of one -- these passes are latency-bound on the bytes in flight), the forward 8.
//   MODE 0: a1[m, :] = relu(scale (W x_m) + shift)                                             (forward)
//   MODE 1: s1 += g, s2 += g xhat,  g = dy (z scale + shift > 0), xhat = (z - mean) rstd        (backward statistics)
//   MODE 2: dz = scale (g - k1 - xhat k2) = g A + z B + D  [k = s / M; eval: dz = scale g];  dW += dz x^T
template <int CPT>
__device__ __forceinline__ void emb_load(const __nv_bfloat16* p, float (&f)[CPT]) {
  if (CPT == 8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < CPT / 2; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  } else {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const uint32_t w[2] = {v.x, v.y};
#pragma unroll
    for (int i = 0; i < CPT / 2; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}

template <int MODE, int CPT>
__global__ void __launch_bounds__(EMB_THREADS, MODE == 0 ? 2 : 3)
embed_l1_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ mean,
                const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                const __nv_bfloat16* __restrict__ dy, const double* __restrict__ s1_in, const double* __restrict__ s2_in,
                double inv_m, long long M, int N, int K, int rows_per_cta, __nv_bfloat16* __restrict__ y,
                double* __restrict__ s1_out, double* __restrict__ s2_out, float* __restrict__ dW) {
  __shared__ float sh_red[MODE == 0 ? 1 : EMB_THREADS * CPT];
  const int tpr = N / CPT, rl = EMB_THREADS / tpr;
  const int cl = threadIdx.x % tpr, r_lane = threadIdx.x / tpr;
  const int n0 = cl * CPT;
  // per-channel constants: sc / sf decide the ReLU mask; MODE 1: (mu, rs) for xhat; MODE 2: dz = g sc + z cb + cd
  float w[CPT][EMB_MAXK], sc[CPT], sf[CPT], ca[CPT], cb[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
#pragma unroll
    for (int c = 0; c < EMB_MAXK; ++c) w[i][c] = c < K ? __ldg(W + (size_t)(n0 + i) * K + c) : 0.f;
    sc[i] = __ldg(scale + n0 + i);
    sf[i] = __ldg(shift + n0 + i);
    ca[i] = cb[i] = 0.f;
    if (MODE == 1) {
      ca[i] = __ldg(mean + n0 + i);
      cb[i] = __ldg(rstd + n0 + i);
    }
    if (MODE == 2 && s1_in != nullptr) {
      const float k1 = (float)(s1_in[n0 + i] * inv_m), k2 = (float)(s2_in[n0 + i] * inv_m);
      const float mu = __ldg(mean + n0 + i), rs = __ldg(rstd + n0 + i);
      ca[i] = -rs * k2 * sc[i];                       // coefficient of z
      cb[i] = (-k1 + mu * rs * k2) * sc[i];           // constant
    }
  }
  float acc[MODE == 2 ? EMB_MAXK : 2][CPT];
#pragma unroll
  for (int q = 0; q < (MODE == 2 ? EMB_MAXK : 2); ++q)
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[q][i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
#pragma unroll 2
  for (long long r = r0 + r_lane; r < r1; r += rl) {
    float xv[EMB_MAXK] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < K; ++c) xv[c] = __ldg(x + (size_t)r * K + c);
    float g[CPT];
    if (MODE != 0) emb_load<CPT>(dy + (size_t)r * N + n0, g);
    float z[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      float t = 0.f;
#pragma unroll
      for (int c = 0; c < EMB_MAXK; ++c) t = fmaf(xv[c], w[i][c], t);
      z[i] = t;
    }
    if (MODE == 0) {
      float o[8];
#pragma unroll
      for (int i = 0; i < CPT; ++i) o[i] = fmaxf(fmaf(z[i], sc[i], sf[i]), 0.f);
      emb_store8(y + (size_t)r * N + n0, o);
    } else {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        const bool on = fmaf(z[i], sc[i], sf[i]) > 0.f;
        const float gi = on ? g[i] : 0.f;
        if (MODE == 1) {
          acc[0][i] += gi;
          acc[1][i] = fmaf(gi, (z[i] - ca[i]) * cb[i], acc[1][i]);
        } else {
          const float dz = fmaf(gi, sc[i], fmaf(z[i], ca[i], cb[i]));
#pragma unroll
          for (int c = 0; c < EMB_MAXK; ++c) acc[c][i] = fmaf(dz, xv[c], acc[c][i]);
        }
      }
    }
  }
  if (MODE == 0) return;
  // combine the rl row lanes of the CTA that own the same channels (fixed order), then one atomic per output per CTA
#pragma unroll
  for (int q = 0; q < (MODE == 2 ? EMB_MAXK : 2); ++q) {
    if (MODE == 2 && q >= K) break;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CPT; ++i) sh_red[threadIdx.x * CPT + i] = acc[q][i];
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += EMB_THREADS) {
      const int t = n / CPT, i = n % CPT;
      float tot = 0.f;
      for (int l = 0; l < rl; ++l) tot += sh_red[(l * tpr + t) * CPT + i];
      if (MODE == 1) atomicAdd((q == 0 ? s1_out : s2_out) + n, (double)tot);
      else atomicAdd(dW + (size_t)n * K + q, tot);
    }
  }
}

// ---- the two backward passes behind a bulk-TMA ring (the structure of stream_bn.cu): the per-thread-load versions above
// keep ~12 KB in flight per SM and ran at 0.75 TB/s; here a producer lane streams 64-row tiles of dy (8 KB) and of the
// coordinates (64 K floats) into an 8-stage ring, 8 consumer warps read them from shared memory.  Needs M % 4 == 0 (tile
// byte counts are multiples of 16) and 16-byte aligned operands, N = 64.
#ifndef P2R_HOST_EMULATION
#define EMBS_ROWS 64
#define EMBS_CONSUMERS 256
#define EMBS_THREADS (EMBS_CONSUMERS + 32)
#define EMBS_STAGES 8

template <int MODE, int K>
__global__ void __launch_bounds__(EMBS_THREADS, 2)
embed_l1_stream_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ mean,
                       const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                       const __nv_bfloat16* __restrict__ dy, const double* __restrict__ s1_in,
                       const double* __restrict__ s2_in, double inv_m, long long M, double* __restrict__ s1_out,
                       double* __restrict__ s2_out, float* __restrict__ dW) {
  extern __shared__ __align__(128) uint8_t embs_smem[];
  constexpr int DY_BYTES = EMBS_ROWS * 128, X_BYTES = EMBS_ROWS * K * 4, STAGE_BYTES = DY_BYTES + ((X_BYTES + 127) / 128) * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(embs_smem + EMBS_STAGES * STAGE_BYTES);
  uint64_t* empty = full + EMBS_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (M + EMBS_ROWS - 1) / EMBS_ROWS;
  if (tid == 0) {
    for (int s = 0; s < EMBS_STAGES; ++s) {
      p2r_mbar_init(full + s, 1);
      p2r_mbar_init(empty + s, EMBS_CONSUMERS / 32);
    }
    p2r_fence_mbar_init();
  }
  __syncthreads();
  if (warp == EMBS_CONSUMERS / 32) {
    if (lane == 0) {
      int it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % EMBS_STAGES;
        p2r_mbar_wait(empty + s, ((uint32_t)(it / EMBS_STAGES) & 1u) ^ 1u);
        const long long r0 = tile * EMBS_ROWS;
        const uint32_t nr = (uint32_t)min((long long)EMBS_ROWS, M - r0);
        p2r_mbar_expect_tx(full + s, nr * 128u + nr * (uint32_t)(K * 4));
        p2r_bulk_g2s(embs_smem + s * STAGE_BYTES, dy + r0 * 64, nr * 128u, full + s);
        p2r_bulk_g2s(embs_smem + s * STAGE_BYTES + DY_BYTES, x + r0 * K, nr * (uint32_t)(K * 4), full + s);
      }
    }
    return;
  }
  const int cv = tid & 7, rl = tid >> 3, n0 = cv * 8;
  float w[8][K], sc[8], sf[8], ca[8], cb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int c = 0; c < K; ++c) w[i][c] = __ldg(W + (size_t)(n0 + i) * K + c);
    sc[i] = __ldg(scale + n0 + i);
    sf[i] = __ldg(shift + n0 + i);
    ca[i] = cb[i] = 0.f;
    if (MODE == 1) {
      ca[i] = __ldg(mean + n0 + i);
      cb[i] = __ldg(rstd + n0 + i);
    }
    if (MODE == 2 && s1_in != nullptr) {
      const float k1 = (float)(s1_in[n0 + i] * inv_m), k2 = (float)(s2_in[n0 + i] * inv_m);
      const float mu = __ldg(mean + n0 + i), rs = __ldg(rstd + n0 + i);
      ca[i] = -rs * k2 * sc[i];
      cb[i] = (-k1 + mu * rs * k2) * sc[i];
    }
  }
  constexpr int NACC = MODE == 2 ? K : 2;
  float acc[NACC][8];
#pragma unroll
  for (int q = 0; q < NACC; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % EMBS_STAGES;
    p2r_mbar_wait(full + s, (uint32_t)(it / EMBS_STAGES) & 1u);
    const int rows_here = (int)min((long long)EMBS_ROWS, M - tile * EMBS_ROWS);
    const uint8_t* st = embs_smem + s * STAGE_BYTES;
    const float* xs = reinterpret_cast<const float*>(st + DY_BYTES);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = rl + 32 * h;
      if (r < rows_here) {
        float g[8];
        {
          const uint4 v = *reinterpret_cast<const uint4*>(st + r * 128 + cv * 16);
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            g[2 * i] = __uint_as_float(u[i] << 16);
            g[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
          }
        }
        float xv[K];
#pragma unroll
        for (int c = 0; c < K; ++c) xv[c] = xs[r * K + c];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float z = 0.f;
#pragma unroll
          for (int c = 0; c < K; ++c) z = fmaf(xv[c], w[i][c], z);
          const float gi = fmaf(z, sc[i], sf[i]) > 0.f ? g[i] : 0.f;
          if (MODE == 1) {
            acc[0][i] += gi;
            acc[1][i] = fmaf(gi, (z - ca[i]) * cb[i], acc[1][i]);
          } else {
            const float dz = fmaf(gi, sc[i], fmaf(z, ca[i], cb[i]));
#pragma unroll
            for (int c = 0; c < K; ++c) acc[c][i] = fmaf(dz, xv[c], acc[c][i]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) p2r_mbar_arrive(empty + s);
  }
  // lanes l, l^8, l^16, l^24 own the same channels: fold them, then combine the 8 warps through shared memory
#pragma unroll
  for (int q = 0; q < NACC; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], 8);
      acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], 16);
    }
  asm volatile("bar.sync 1, 256;" ::: "memory");          // every consumer is past its last tile: the ring is free
  float* red = reinterpret_cast<float*>(embs_smem);       // [NACC][8 warps][64 channels]
  if (lane < 8) {
#pragma unroll
    for (int q = 0; q < NACC; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[(q * 8 + warp) * 64 + n0 + i] = acc[q][i];
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
  for (int e = tid; e < NACC * 64; e += EMBS_CONSUMERS) {
    const int q = e >> 6, c = e & 63;
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) t += red[(q * 8 + w8) * 64 + c];
    if (MODE == 1) atomicAdd((q == 0 ? s1_out : s2_out) + c, (double)t);
    else atomicAdd(dW + (size_t)c * K + q, t);
  }
}

template <int MODE>
static int embs_launch(const float* x, const float* W, const float* mean, const float* rstd, const float* scale,
                       const float* shift, const void* dy, const double* s1_in, const double* s2_in, long long M, int K,
                       double* s1_out, double* s2_out, float* dW, cudaStream_t st) {
  const long long ntiles = (M + EMBS_ROWS - 1) / EMBS_ROWS;
  const int grid = (int)min(ntiles, (long long)P2R_SM_COUNT * 2);
#define EMBS_GO(KK)                                                                                                  \
  do {                                                                                                               \
    constexpr int SMEM = EMBS_STAGES * (EMBS_ROWS * 128 + ((EMBS_ROWS * KK * 4 + 127) / 128) * 128) + 2 * EMBS_STAGES * 8; \
    auto kern = embed_l1_stream_kernel<MODE, KK>;                                                                    \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);                                   \
    kern<<<grid, EMBS_THREADS, SMEM, st>>>(x, W, mean, rstd, scale, shift, (const __nv_bfloat16*)dy, s1_in, s2_in,   \
                                            1.0 / (double)M, M, s1_out, s2_out, dW);                                  \
  } while (0)
  if (K == 3) EMBS_GO(3);
  else if (K == 4) EMBS_GO(4);
  else if (K == 2) EMBS_GO(2);
  else EMBS_GO(1);
#undef EMBS_GO
  return 0;
}

static bool embs_ok(long long M, int N, int K, const void* x, const void* dy) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("P2R_EMBED_STREAM");
    enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return enabled == 1 && N == 64 && M >= 4096 && M % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0;
}
#else
static bool embs_ok(long long, int, int, const void*, const void*) { return false; }
#endif

// Forward for N = 64: the coordinates of 256 rows are staged in shared memory with coalesced loads (the next chunk's loads
// are in flight while this one is computed); the version above kept two rows per thread in flight behind scalar loads and
// wrote at 1.9 TB/s.  One thread = 8 channels of a row, 8 rows per chunk.
template <int K>
__global__ void __launch_bounds__(EMB_THREADS, 4)
embed_l1_fwd64_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ scale,
                      const float* __restrict__ shift, long long M, int rows_per_cta, __nv_bfloat16* __restrict__ y) {
  constexpr int CH = 256;                               // rows per chunk
  __shared__ float sx[2][CH * K];
  const int cl = threadIdx.x & 7, r_lane = threadIdx.x >> 3, n0 = cl * 8;
  float w[8][K], sc[8], sf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int c = 0; c < K; ++c) w[i][c] = __ldg(W + (size_t)(n0 + i) * K + c);
    sc[i] = __ldg(scale + n0 + i);
    sf[i] = __ldg(shift + n0 + i);
  }
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float pre[K];
  auto fetch = [&](long long rc) {                      // x[rc .. rc + CH) is one contiguous run of CH * K floats
    const long long lim = (r1 - rc) * K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int e = threadIdx.x + j * EMB_THREADS;
      pre[j] = e < lim ? __ldg(x + rc * K + e) : 0.f;
    }
  };
  if (r0 < r1) fetch(r0);
  int buf = 0;
  for (long long rc = r0; rc < r1; rc += CH, buf ^= 1) {
#pragma unroll
    for (int j = 0; j < K; ++j) sx[buf][threadIdx.x + j * EMB_THREADS] = pre[j];
    __syncthreads();                                    // (two buffers: the previous chunk's readers are one barrier behind)
    if (rc + CH < r1) fetch(rc + CH);
#pragma unroll
    for (int j = 0; j < CH / 32; ++j) {
      const int rr = j * 32 + r_lane;
      const long long r = rc + rr;
      if (r < r1) {
        float xv[K];
#pragma unroll
        for (int c = 0; c < K; ++c) xv[c] = sx[buf][rr * K + c];
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = 0.f;
#pragma unroll
          for (int c = 0; c < K; ++c) t = fmaf(xv[c], w[i][c], t);
          o[i] = fmaxf(fmaf(t, sc[i], sf[i]), 0.f);
        }
        emb_store8(y + (size_t)r * 64 + n0, o);
      }
    }
  }
}

static int emb_grid(long long M, int* rows_per_cta) {
  const long long ctas = (long long)P2R_SM_COUNT * 8;
  long long target = (M + ctas - 1) / ctas;
  if (target < 64) target = 64;
  *rows_per_cta = (int)target;
  return (int)((M + target - 1) / target);
}

static int emb_check(long long M, int N, int K, const char* where) {
  if (!(M >= 0 && K >= 1 && K <= EMB_MAXK && N >= 8 && N % 8 == 0 && 2048 % N == 0 && EMB_THREADS % (N / 4) == 0)) {
    p2r_set_last_error(where, -1);
    return -1;
  }
  return 0;
}

extern "C" int p2r_embed_l1_fwd(const float* x, const float* W, const float* scale, const float* shift, long long M, int N,
                                int K, void* y, void* stream) {
  if (emb_check(M, N, K, "p2r_embed_l1_fwd: bad argument")) return -1;
  if (M == 0) return 0;
  int rpc;
  const int grid = emb_grid(M, &rpc);
#ifndef P2R_HOST_EMULATION
  if (N == 64 && K >= 1 && K <= 4) {
    rpc = (rpc + 255) / 256 * 256;                      // whole chunks per CTA
    const int g64 = (int)((M + rpc - 1) / rpc);
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 3) embed_l1_fwd64_kernel<3><<<g64, EMB_THREADS, 0, st>>>(x, W, scale, shift, M, rpc, (__nv_bfloat16*)y);
    else if (K == 4) embed_l1_fwd64_kernel<4><<<g64, EMB_THREADS, 0, st>>>(x, W, scale, shift, M, rpc, (__nv_bfloat16*)y);
    else if (K == 2) embed_l1_fwd64_kernel<2><<<g64, EMB_THREADS, 0, st>>>(x, W, scale, shift, M, rpc, (__nv_bfloat16*)y);
    else embed_l1_fwd64_kernel<1><<<g64, EMB_THREADS, 0, st>>>(x, W, scale, shift, M, rpc, (__nv_bfloat16*)y);
    P2R_RETURN_LAUNCH("p2r_embed_l1_fwd");
  }
#endif
  auto kern = embed_l1_kernel<0, 8>;
  P2R_LAUNCH(kern, grid, EMB_THREADS, 0, (cudaStream_t)stream, x, W, (const float*)nullptr, (const float*)nullptr, scale,
             shift, (const __nv_bfloat16*)nullptr, (const double*)nullptr, (const double*)nullptr, 0.0, M, N, K, rpc,
             (__nv_bfloat16*)y, (double*)nullptr, (double*)nullptr, (float*)nullptr);
  P2R_RETURN_LAUNCH("p2r_embed_l1_fwd");
}

// s1, s2: double[N], zero-filled by the caller.
extern "C" int p2r_embed_l1_bwd_stats(const void* dy, const float* x, const float* W, const float* mean, const float* rstd,
                                      const float* scale, const float* shift, long long M, int N, int K, double* s1,
                                      double* s2, void* stream) {
  if (emb_check(M, N, K, "p2r_embed_l1_bwd_stats: bad argument")) return -1;
  if (M == 0) return 0;
#ifndef P2R_HOST_EMULATION
  if (embs_ok(M, N, K, x, dy)) {
    embs_launch<1>(x, W, mean, rstd, scale, shift, dy, nullptr, nullptr, M, K, s1, s2, nullptr, (cudaStream_t)stream);
    P2R_RETURN_LAUNCH("p2r_embed_l1_bwd_stats");
  }
#endif
  int rpc;
  const int grid = emb_grid(M, &rpc);
  auto kern = embed_l1_kernel<1, 4>;
  P2R_LAUNCH(kern, grid, EMB_THREADS, 0, (cudaStream_t)stream, x, W, mean, rstd, scale, shift, (const __nv_bfloat16*)dy,
             (const double*)nullptr, (const double*)nullptr, 0.0, M, N, K, rpc, (__nv_bfloat16*)nullptr, s1, s2,
             (float*)nullptr);
  P2R_RETURN_LAUNCH("p2r_embed_l1_bwd_stats");
}

// dW: float[N, K], zero-filled by the caller.  s1 / s2 = the sums of p2r_embed_l1_bwd_stats (NULL: eval-mode BatchNorm).
extern "C" int p2r_embed_l1_bwd_dw(const void* dy, const float* x, const float* W, const float* mean, const float* rstd,
                                   const float* scale, const float* shift, const double* s1, const double* s2,
                                   long long M, int N, int K, float* dW, void* stream) {
  if (emb_check(M, N, K, "p2r_embed_l1_bwd_dw: bad argument")) return -1;
  if (M == 0) return 0;
#ifndef P2R_HOST_EMULATION
  if (embs_ok(M, N, K, x, dy)) {
    embs_launch<2>(x, W, mean, rstd, scale, shift, dy, s1, s2, M, K, nullptr, nullptr, dW, (cudaStream_t)stream);
    P2R_RETURN_LAUNCH("p2r_embed_l1_bwd_dw");
  }
#endif
  int rpc;
  const int grid = emb_grid(M, &rpc);
  auto kern = embed_l1_kernel<2, 4>;
  P2R_LAUNCH(kern, grid, EMB_THREADS, 0, (cudaStream_t)stream, x, W, mean, rstd, scale, shift, (const __nv_bfloat16*)dy, s1,
             s2, 1.0 / (double)M, M, N, K, rpc, (__nv_bfloat16*)nullptr, (double*)nullptr, (double*)nullptr, dW);
  P2R_RETURN_LAUNCH("p2r_embed_l1_bwd_dw");
}
