// dense_ops.cu -- per-point dense layers of the P2RNet hot path in CHANNEL-LAST layout
// (rows = points / frames, columns = channels), so that every 1x1 conv of the reference
// (models/p2rnet/modules/stgcn.py:45-67, stgcn_layers.py:50-56,402-414, vote_center.py:28-32,
//  proposal_net.py:78-94, pointnet2_modules.py:9-19) is a row-major GEMM and every BatchNorm is a
// column statistic.  This file holds the fp32-exact SIMT GEMM (parity mode, and the small-K layers of
// the bf16 mode), the training-mode BatchNorm forward/backward, the temporal unfold/fold, the
// channel-last grouping + max-pool of the set-abstraction layer, and small fused elementwise pieces.
// The bf16 tensor-core GEMM (tcgen05 + TMA) lives in gemm_sm100.cu.
//
// Storage type T is float or __nv_bfloat16; arithmetic is always fp32 (double for BN column sums).
#include "p2r_common.cuh"
#include "stream_bn.cuh"
#include <stdlib.h>

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}


// 16-byte vector access: 4 floats or 8 bf16 per thread per load/store (HBM-bound elementwise kernels)
template <typename T> struct VecN;
template <> struct VecN<float> { static constexpr int N = 4; };
template <> struct VecN<__nv_bfloat16> { static constexpr int N = 8; };
__device__ __forceinline__ void vload(const float* p, float (&f)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void vload(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void vstore(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ void vstore(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(p) = v;
}
template <typename T> static inline bool vec_ok(int C, const void* a, const void* b = nullptr, const void* c = nullptr,
                                                const void* d = nullptr, const void* e = nullptr) {
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return C % VecN<T>::N == 0 && (256 * VecN<T>::N) % C == 0 && al(a) && al(b) && al(c) && al(d) && al(e);
}

// ================================================================================================
// SIMT GEMM:  C[M,N] (+)= op(A)[M,K] . op(B)[K,N]  (+ bias[N]) (ReLU)
//   TRANS_A = false: A is [M,K] row-major (lda);  true: A is [K,M] row-major (lda)
//   TRANS_B = true : B is [N,K] row-major (ldb) -- the nn.Linear / 1x1-conv weight layout;
//             false: B is [K,N] row-major (ldb)
//   gridDim.z > 1 = split-K with fp32 atomicAdd into a zero-filled C (TC must be float, no epilogue).
// 64x64 tile, 16-deep k slab, 256 threads, 4x4 micro-tile; fixed summation order along K inside a
// split => deterministic when gridDim.z == 1.
// ================================================================================================
#define SG_BM 64
#define SG_BN 64
#define SG_BK 16
template <typename TA, typename TB, typename TC, bool TRANS_A, bool TRANS_B>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const TA* __restrict__ A, int lda, const TB* __restrict__ B, int ldb,
             TC* __restrict__ C, int ldc, const float* __restrict__ bias, int relu, int accumulate, int k_per_split) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4 (m) x 4 (n)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += SG_BK) {
    // ---- stage A tile: As[k][m]
#pragma unroll
    for (int e = tid; e < SG_BM * SG_BK; e += 256) {
      int m, k;
      if (!TRANS_A) { m = e / SG_BK; k = e % SG_BK; }   // k fastest: coalesced along a row of A
      else          { k = e / SG_BM; m = e % SG_BM; }   // m fastest: coalesced along a row of A^T storage
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = TRANS_A ? ldf<TA>(A + (size_t)gk * lda + gm) : ldf<TA>(A + (size_t)gm * lda + gk);
      As[k][m] = v;
    }
    // ---- stage B tile: Bs[k][n]
#pragma unroll
    for (int e = tid; e < SG_BN * SG_BK; e += 256) {
      int n, k;
      if (TRANS_B) { n = e / SG_BK; k = e % SG_BK; }
      else         { k = e / SG_BN; n = e % SG_BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = TRANS_B ? ldf<TB>(B + (size_t)gn * ldb + gk) : ldf<TB>(B + (size_t)gk * ldb + gn);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (gridDim.z > 1) {
        atomicAdd(reinterpret_cast<float*>(C) + (size_t)gm * ldc + gn, v);
      } else {
        if (bias) v += __ldg(bias + gn);
        if (accumulate) v += ldf<TC>(C + (size_t)gm * ldc + gn);
        if (relu) v = fmaxf(v, 0.f);
        stf<TC>(C + (size_t)gm * ldc + gn, v);
      }
    }
  }
}

template <typename TA, typename TB, typename TC>
static int launch_sgemm(int M, int N, int K, const void* A, int lda, int trans_a, const void* B, int ldb, int trans_b,
                        void* C, int ldc, const float* bias, int relu, int accumulate, int splits, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  int kps = K;
  if (splits > 1) {
    kps = ((K + splits - 1) / splits + SG_BK - 1) / SG_BK * SG_BK;
    splits = (K + kps - 1) / kps;
  } else splits = 1;
  dim3 grid(p2r_ceil_div(N, SG_BN), p2r_ceil_div(M, SG_BM), splits);
#define SG_LAUNCH(TA_, TB_)                                                                                   \
  sgemm_kernel<TA, TB, TC, TA_, TB_><<<grid, 256, 0, st>>>(M, N, K, (const TA*)A, lda, (const TB*)B, ldb,    \
                                                            (TC*)C, ldc, bias, relu, accumulate, kps)
  if (!trans_a && trans_b) SG_LAUNCH(false, true);
  else if (!trans_a && !trans_b) SG_LAUNCH(false, false);
  else if (trans_a && !trans_b) SG_LAUNCH(true, false);
  else SG_LAUNCH(true, true);
#undef SG_LAUNCH
  P2R_RETURN_LAUNCH("p2r_sgemm");
}

// dtype codes: 0 = float32, 1 = bfloat16
extern "C" int p2r_sgemm(int M, int N, int K, const void* A, int lda, int trans_a, int a_dtype, const void* B, int ldb,
                         int trans_b, int b_dtype, void* C, int ldc, int c_dtype, const float* bias, int relu,
                         int accumulate, int splits, void* stream) {
  P2R_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "p2r_sgemm");
  P2R_CHECK_ARG(!(splits > 1 && (c_dtype != 0 || bias || relu)), "p2r_sgemm (split-K needs fp32 C and no epilogue)");
  cudaStream_t st = (cudaStream_t)stream;
  const int code = a_dtype * 4 + b_dtype * 2 + c_dtype;
  switch (code) {
    case 0: return launch_sgemm<float, float, float>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    case 1: return launch_sgemm<float, float, __nv_bfloat16>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    case 2: return launch_sgemm<float, __nv_bfloat16, float>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    case 3: return launch_sgemm<float, __nv_bfloat16, __nv_bfloat16>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    case 4: return launch_sgemm<__nv_bfloat16, float, float>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    case 5: return launch_sgemm<__nv_bfloat16, float, __nv_bfloat16>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    case 6: return launch_sgemm<__nv_bfloat16, __nv_bfloat16, float>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
    default: return launch_sgemm<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(M, N, K, A, lda, trans_a, B, ldb, trans_b, C, ldc, bias, relu, accumulate, splits, st);
  }
}

// ================================================================================================
// column reductions over rows of a [M,C] matrix (BatchNorm statistics, bias gradients)
//   mode 0: s1 = sum x,            s2 = sum x*x
//   mode 1: s1 = sum dz,           s2 = sum dz * (x - mean) * rstd      dz = dy * (y > 0) if relu else dy
// Thread layout: 256 threads = RL row lanes x CP channels (CP = min(C,256), C | 256 or 256 | C).
// Per-thread fp32 partials over <= 64 rows, then double atomics into s1/s2 (zero-filled by caller).
// ================================================================================================
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
colreduce_kernel(long long M, int C, const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ y,
                 const float* __restrict__ mean, const float* __restrict__ rstd, int relu, int rows_per_cta,
                 double* __restrict__ s1, double* __restrict__ s2, const float* __restrict__ scale = nullptr,
                 const float* __restrict__ shift = nullptr) {
  __shared__ float sh1[256], sh2[256];
  const int cp = C < 256 ? C : 256;
  const int rl = 256 / cp;
  const int cl = threadIdx.x % cp, r_lane = threadIdx.x / cp;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(M, r0 + rows_per_cta);
  for (int cbase = 0; cbase < C; cbase += cp) {
    const int c = cbase + cl;
    float a1 = 0.f, a2 = 0.f;
    float mu = 0.f, rs = 1.f;
    if (MODE == 1 && mean != nullptr) { mu = __ldg(mean + c); rs = __ldg(rstd + c); }
    for (long long r = r0 + r_lane; r < r1; r += rl) {
      const size_t o = (size_t)r * C + c;
      if (MODE == 0) {
        const float v = ldf<T>(x + o);
        a1 += v;
        a2 = fmaf(v, v, a2);
      } else {
        float dz = ldf<T>(dy + o);
        if (relu == 1 && !(ldf<T>(y + o) > 0.f)) dz = 0.f;
        if (relu == 2 && !(fmaf(ldf<T>(x + o), __ldg(scale + c), __ldg(shift + c)) > 0.f)) dz = 0.f;
        a1 += dz;
        if (x) a2 = fmaf(dz, (ldf<T>(x + o) - mu) * rs, a2);
      }
    }
    sh1[threadIdx.x] = a1;
    sh2[threadIdx.x] = a2;
    __syncthreads();
    if (r_lane == 0) {
      double t1 = 0.0, t2 = 0.0;
      for (int l = 0; l < rl; ++l) { t1 += (double)sh1[l * cp + cl]; t2 += (double)sh2[l * cp + cl]; }
      atomicAdd(s1 + c, t1);
      if (s2) atomicAdd(s2 + c, t2);
    }
    __syncthreads();
  }
}

// Vectorised variant: thread = VEC consecutive channels of one row lane; C % VEC == 0, 256 % (C/VEC) == 0.
// MODE 1 accumulates b = sum dz*x and folds s2 = rstd * (b - mean * s1) when the CTA's partials are combined (in
// double), so mean / rstd never sit in registers inside the streaming loop; two rows are in flight per iteration.
template <typename T, int MODE>
__global__ void __launch_bounds__(256, 3)
colreduce_vec_kernel(long long M, int C, const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ y,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int relu, int rows_per_cta,
                     double* __restrict__ s1, double* __restrict__ s2, const float* __restrict__ scale = nullptr,
                     const float* __restrict__ shift = nullptr) {
  constexpr int V = VecN<T>::N;
  __shared__ float sh1[256 * V], sh2[256 * V];
  const int tpr = C / V;            // threads per row
  const int rl = 256 / tpr;         // row lanes
  const int cl = threadIdx.x % tpr, r_lane = threadIdx.x / tpr;
  const int c0 = cl * V;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(M, r0 + rows_per_cta);
  float a1[V], a2[V], sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    a1[i] = a2[i] = 0.f;
    sc[i] = sh[i] = 0.f;
    if (MODE == 1 && relu == 2) { sc[i] = __ldg(scale + c0 + i); sh[i] = __ldg(shift + c0 + i); }
  }
  auto body = [&](const float (&p)[V], const float (&q)[V], const float (&w)[V]) {
    // MODE 0: p = x.   MODE 1: p = dy, q = x (if any), w = y (relu == 1)
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) { a1[i] += p[i]; a2[i] = fmaf(p[i], p[i], a2[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float dz = p[i];
        if (relu == 1 && !(w[i] > 0.f)) dz = 0.f;
        if (relu == 2 && !(fmaf(q[i], sc[i], sh[i]) > 0.f)) dz = 0.f;
        a1[i] += dz;
        a2[i] = fmaf(dz, q[i], a2[i]);
      }
    }
  };
  const bool need_x = MODE == 1 && x != nullptr;
  long long r = r0 + r_lane;
  for (; r + rl < r1; r += 2 * rl) {   // two rows per iteration, all loads issued before use
    const size_t o0 = (size_t)r * C + c0, o1 = (size_t)(r + rl) * C + c0;
    float p0[V], p1[V], q0[V], q1[V], w0[V], w1[V];
    if (MODE == 0) { vload(x + o0, p0); vload(x + o1, p1); }
    else { vload(dy + o0, p0); vload(dy + o1, p1); }
    if (need_x) { vload(x + o0, q0); vload(x + o1, q1); }
    else {
#pragma unroll
      for (int i = 0; i < V; ++i) q0[i] = q1[i] = 0.f;
    }
    if (MODE == 1 && relu == 1) { vload(y + o0, w0); vload(y + o1, w1); }
    else {
#pragma unroll
      for (int i = 0; i < V; ++i) w0[i] = w1[i] = 1.f;
    }
    body(p0, q0, w0);
    body(p1, q1, w1);
  }
  for (; r < r1; r += rl) {
    const size_t o0 = (size_t)r * C + c0;
    float p0[V], q0[V], w0[V];
    if (MODE == 0) vload(x + o0, p0); else vload(dy + o0, p0);
#pragma unroll
    for (int i = 0; i < V; ++i) { q0[i] = 0.f; w0[i] = 1.f; }
    if (need_x) vload(x + o0, q0);
    if (MODE == 1 && relu == 1) vload(y + o0, w0);
    body(p0, q0, w0);
  }
#pragma unroll
  for (int i = 0; i < V; ++i) { sh1[threadIdx.x * V + i] = a1[i]; sh2[threadIdx.x * V + i] = a2[i]; }
  __syncthreads();
  // one thread per channel finishes the reduction over the row lanes
  for (int c = threadIdx.x; c < C; c += 256) {
    const int t = c / V, i = c % V;
    double t1 = 0.0, t2 = 0.0;
    for (int l = 0; l < rl; ++l) { t1 += (double)sh1[(l * tpr + t) * V + i]; t2 += (double)sh2[(l * tpr + t) * V + i]; }
    atomicAdd(s1 + c, t1);
    if (s2) {
      if (MODE == 1 && mean != nullptr) t2 = (double)__ldg(rstd + c) * (t2 - (double)__ldg(mean + c) * t1);
      atomicAdd(s2 + c, t2);
    }
  }
}

static int colreduce_ctas_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("P2R_COLREDUCE_CTAS_PER_SM");  // tuning knob (default 8)
    v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
  }
  return v;
}

static int colreduce_grid(long long M, int* rows_per_cta) {
  // ~8 CTAs per SM, at least 64 rows each
  const long long ctas = (long long)P2R_SM_COUNT * colreduce_ctas_per_sm();
  long long target = (M + ctas - 1) / ctas;
  if (target < 64) target = 64;
  *rows_per_cta = (int)target;
  return (int)((M + target - 1) / target);
}

// sum / sum-of-squares per column. s1, s2: double[C], zero-filled by the caller.
extern "C" int p2r_col_stats(const void* x, int dtype, long long M, int C, double* s1, double* s2, void* stream) {
  P2R_CHECK_ARG(M >= 0 && C > 0 && (C <= 256 ? 256 % C == 0 : C % 256 == 0), "p2r_col_stats");
  if (M == 0) return 0;
  if (p2r_stream_bn_ok(dtype, M, C, x)) return p2r_stream_col_stats(x, M, s1, s2, (cudaStream_t)stream);
  int rpc;
  const int grid = colreduce_grid(M, &rpc);
  if (dtype == 0) {
    if (vec_ok<float>(C, x) && 256 % (C / 4) == 0)
      colreduce_vec_kernel<float, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const float*)x, nullptr, nullptr, nullptr, nullptr, 0, rpc, s1, s2);
    else
      colreduce_kernel<float, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const float*)x, nullptr, nullptr, nullptr, nullptr, 0, rpc, s1, s2);
  } else {
    if (vec_ok<__nv_bfloat16>(C, x) && 256 % (C / 8) == 0)
      colreduce_vec_kernel<__nv_bfloat16, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, nullptr, 0, rpc, s1, s2);
    else
      colreduce_kernel<__nv_bfloat16, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const __nv_bfloat16*)x, nullptr, nullptr, nullptr, nullptr, 0, rpc, s1, s2);
  }
  P2R_RETURN_LAUNCH("p2r_col_stats");
}

// column sums of dz = relu ? dy * (y > 0) : dy for ANY width (the 259-column vote head, 100 mixture weights, 24 box
// parameters): a block takes a run of rows, adds into a shared float table, one double atomic per column and block.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_any_kernel(long long M, int C, int rows_per_block, const T* __restrict__ dy, const T* __restrict__ y,
                  double* __restrict__ s1) {
  // thread t owns columns t, t + 256, ...: consecutive threads read consecutive elements of a row (coalesced), the sum
  // over the block's rows stays in a register, one double atomic per column and block.  (The first version added every
  // element into a shared table with an atomic and an integer modulo: 41 us for 8.5 MB.)
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
#pragma unroll 8
    for (long long r = r0; r < r1; ++r) {
      float v = ldf<T>(dy + r * C + c);
      if (y != nullptr && !(ldf<T>(y + r * C + c) > 0.f)) v = 0.f;
      acc += v;
    }
    atomicAdd(s1 + c, (double)acc);
  }
}

// backward column sums: s1 = sum dz, s2 = sum dz*xhat (s2/x/mean/rstd may be NULL -> only s1, e.g. a bias grad)
extern "C" int p2r_col_bwd_stats(const void* dy, const void* x, const void* y, int dtype, long long M, int C,
                                 const float* mean, const float* rstd, int relu, double* s1, double* s2,
                                 const float* scale, const float* shift, void* stream) {
  const bool any_width = x == nullptr && s2 == nullptr && (relu == 0 || relu == 1) && C <= 8192;   // a bias gradient
  P2R_CHECK_ARG(M >= 0 && C > 0 && (any_width || (C <= 256 ? 256 % C == 0 : C % 256 == 0)), "p2r_col_bwd_stats");
  P2R_CHECK_ARG(!(relu == 1 && y == nullptr), "p2r_col_bwd_stats (relu = 1 needs y)");
  if (M > 0 && any_width && !(C <= 256 ? 256 % C == 0 : C % 256 == 0)) {
    const int rpb = (int)max(16LL, min(256LL, (M + P2R_SM_COUNT * 4 - 1) / (P2R_SM_COUNT * 4)));
    const int grid = (int)((M + rpb - 1) / rpb);
    if (dtype == 0)
      colsum_any_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, rpb, (const float*)dy, relu ? (const float*)y : nullptr, s1);
    else
      colsum_any_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, rpb, (const __nv_bfloat16*)dy, relu ? (const __nv_bfloat16*)y : nullptr, s1);
    P2R_RETURN_LAUNCH("p2r_col_bwd_stats");
  }
  P2R_CHECK_ARG(!(relu == 2 && (x == nullptr || scale == nullptr || shift == nullptr)), "p2r_col_bwd_stats (relu = 2 needs x, scale, shift)");
  if (M == 0) return 0;
  if ((x != nullptr || relu == 0) && (s2 == nullptr || x != nullptr) && p2r_stream_bn_ok(dtype, M, C, dy, x, relu == 3 ? nullptr : y))
    return p2r_stream_col_bwd_stats(dy, x, (relu == 1 || relu == 3) ? y : nullptr, M, mean, rstd, relu, s1, s2, scale, shift, (cudaStream_t)stream);
  P2R_CHECK_ARG(relu != 3, "p2r_col_bwd_stats (relu = 3, the bit mask, needs p2r_stream_bn_supported(dtype, M, C))");
  int rpc;
  const int grid = colreduce_grid(M, &rpc);
  if (dtype == 0) {
    if (vec_ok<float>(C, x, dy, y) && 256 % (C / 4) == 0)
      colreduce_vec_kernel<float, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const float*)x, (const float*)dy, (const float*)y, mean, rstd, relu, rpc, s1, s2, scale, shift);
    else
      colreduce_kernel<float, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const float*)x, (const float*)dy, (const float*)y, mean, rstd, relu, rpc, s1, s2, scale, shift);
  } else {
    if (vec_ok<__nv_bfloat16>(C, x, dy, y) && 256 % (C / 8) == 0)
      colreduce_vec_kernel<__nv_bfloat16, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, mean, rstd, relu, rpc, s1, s2, scale, shift);
    else
      colreduce_kernel<__nv_bfloat16, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, mean, rstd, relu, rpc, s1, s2, scale, shift);
  }
  P2R_RETURN_LAUNCH("p2r_col_bwd_stats");
}

// BatchNorm finalize (training): mean/var from the column sums, running-stat update exactly like
// torch.nn.BatchNorm (momentum, unbiased variance for the running estimate), fused affine scale/shift.
__global__ void bn_finalize_kernel(int C, double inv_m, double unbias, int copies, long long copy_stride,
                                   const double* __restrict__ s1,
                                   const double* __restrict__ s2, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                                   float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // the partial sums sixteen at a time: all loads first, then the adds in the same order as a plain loop (same rounding).
  // With the add right behind each load the 16 copies of the fused GEMM statistics were 16 dependent L2 round trips --
  // most of this kernel's 5 us, 21 times on the forward's critical path.
  double t1 = 0.0, t2 = 0.0;
  for (int k0 = 0; k0 < copies; k0 += 16) {
    double a1[16], a2[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const bool ok = k0 + k < copies;
      a1[k] = ok ? s1[(k0 + k) * copy_stride + c] : 0.0;
      a2[k] = ok ? s2[(k0 + k) * copy_stride + c] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) { t1 += a1[k]; t2 += a2[k]; }
  }
  const double mu = t1 * inv_m;
  double var = t2 * inv_m - mu * mu;
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

// s1 / s2 may be given as `copies` partial sums `copy_stride` doubles apart (the fused GEMM-epilogue statistics).
extern "C" int p2r_bn_finalize(int C, long long M, const double* s1, const double* s2, int copies,
                               long long copy_stride, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               float* mean, float* rstd, float* scale, float* shift, void* stream) {
  P2R_CHECK_ARG(C > 0 && M > 0 && copies >= 1, "p2r_bn_finalize");
  const double unbias = M > 1 ? (double)M / (double)(M - 1) : 1.0;
  bn_finalize_kernel<<<p2r_ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(C, 1.0 / (double)M, unbias, copies, copy_stride, s1, s2, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale, shift);
  P2R_RETURN_LAUNCH("p2r_bn_finalize");
}

// y = x*scale[c] + shift[c] (+ residual) (ReLU).  Vectorised over 4 channels when C % 4 == 0.
template <typename T>
__global__ void __launch_bounds__(256)
affine_act_kernel(long long total, int C, const T* __restrict__ x, const float* __restrict__ scale,
                  const float* __restrict__ shift, const T* __restrict__ residual, int relu, T* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    float v = fmaf(ldf<T>(x + e), __ldg(scale + c), __ldg(shift + c));
    if (residual) v += ldf<T>(residual + e);
    if (relu) v = fmaxf(v, 0.f);
    stf<T>(y + e, v);
  }
}

// Vectorised variants require (256 * VEC) % C == 0, so a thread's channel offset is the same in every iteration of
// the grid-stride loop and the per-channel coefficients live in registers (no per-element parameter loads).
// Grid of the vectorised elementwise BatchNorm kernels: every thread first loads the per-channel constants of its 4 / 8
// columns (up to 48 scalar loads + double arithmetic), so a thread should own several vectors -- with one vector per thread
// the 16 384 x 256 passes of the vote MLP took 33 us for 24 MB.
static inline int vec_grid(long long nvec) {
  const long long blocks = (nvec + 255) / 256;
  const long long cap = (long long)P2R_SM_COUNT * 4;
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template <typename T>
__global__ void __launch_bounds__(256)
affine_act_vec_kernel(long long nvec, int C, const T* __restrict__ x, const float* __restrict__ scale,
                      const float* __restrict__ shift, const T* __restrict__ residual, int relu, T* __restrict__ y) {
  constexpr int V = VecN<T>::N;
  const int c0 = (threadIdx.x * V) % C;
  float sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { sc[i] = __ldg(scale + c0 + i); sh[i] = __ldg(shift + c0 + i); }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nvec; e += stride) {
    const size_t o = (size_t)e * V;
    float v[V], r[V];
    vload(x + o, v);
    if (residual) vload(residual + o, r);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float t = fmaf(v[i], sc[i], sh[i]);
      if (residual) t += r[i];
      v[i] = relu ? fmaxf(t, 0.f) : t;
    }
    vstore(y + o, v);
  }
}

extern "C" int p2r_affine_act(const void* x, int dtype, long long M, int C, const float* scale, const float* shift,
                              const void* residual, int relu, void* y, unsigned char* relu_mask, void* stream) {
  P2R_CHECK_ARG(M >= 0 && C > 0, "p2r_affine_act");
  const long long total = M * C;
  if (total == 0) return 0;
  if (p2r_stream_bn_ok(dtype, M, C, x, residual, y))
    return p2r_stream_affine_act(x, M, scale, shift, residual, relu, y, relu_mask, (cudaStream_t)stream);
  P2R_CHECK_ARG(relu_mask == nullptr, "p2r_affine_act (the ReLU bit mask needs p2r_stream_bn_supported(dtype, M, C))");
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (total + 255) / 256);
  if (dtype == 0 && vec_ok<float>(C, x, residual, y))
    affine_act_vec_kernel<float><<<vec_grid(total / 4), 256, 0, (cudaStream_t)stream>>>(total / 4, C, (const float*)x, scale, shift, (const float*)residual, relu, (float*)y);
  else if (dtype == 1 && vec_ok<__nv_bfloat16>(C, x, residual, y))
    affine_act_vec_kernel<__nv_bfloat16><<<vec_grid(total / 8), 256, 0, (cudaStream_t)stream>>>(total / 8, C, (const __nv_bfloat16*)x, scale, shift, (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)y);
  else if (dtype == 0)
    affine_act_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(total, C, (const float*)x, scale, shift, (const float*)residual, relu, (float*)y);
  else
    affine_act_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(total, C, (const __nv_bfloat16*)x, scale, shift, (const __nv_bfloat16*)residual, relu, (__nv_bfloat16*)y);
  P2R_RETURN_LAUNCH("p2r_affine_act");
}

// BatchNorm(+ReLU)(+residual) backward, elementwise part:
//   dz = relu ? dy * (y > 0) : dy
//   training: dx = scale[c] * (dz - s1[c]/M - xhat * s2[c]/M)      eval (s1 == NULL): dx = scale[c] * dz
//   dres (optional) = dz
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(long long total, int C, double inv_m, const T* __restrict__ dy, const T* __restrict__ x,
                    const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ scale, const double* __restrict__ s1, const double* __restrict__ s2,
                    int relu, T* __restrict__ dx, T* __restrict__ dres, const float* __restrict__ shift = nullptr) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int c = (int)(e % C);
    float dz = ldf<T>(dy + e);
    if (relu == 1 && !(ldf<T>(y + e) > 0.f)) dz = 0.f;
    if (relu == 2 && !(fmaf(ldf<T>(x + e), __ldg(scale + c), __ldg(shift + c)) > 0.f)) dz = 0.f;
    if (dres) stf<T>(dres + e, dz);
    float g = dz;
    if (s1) {
      const float xh = (ldf<T>(x + e) - __ldg(mean + c)) * __ldg(rstd + c);
      g = dz - (float)(s1[c] * inv_m) - xh * (float)(s2[c] * inv_m);
    }
    stf<T>(dx + e, g * __ldg(scale + c));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(long long nvec, int C, double inv_m, const T* __restrict__ dy, const T* __restrict__ x,
                        const T* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                        const float* __restrict__ scale, const double* __restrict__ s1, const double* __restrict__ s2,
                        int relu, T* __restrict__ dx, T* __restrict__ dres, const float* __restrict__ shift = nullptr) {
  constexpr int V = VecN<T>::N;
  const int c0 = (threadIdx.x * V) % C;
  // dx = scale * (dz - k1 - xhat * k2),  xhat = (x - mean) * rstd   ==>   dx = dz * A + x * B + D
  float A[V], Bc[V], D[V], sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = c0 + i;
    sc[i] = __ldg(scale + c);
    sh[i] = shift ? __ldg(shift + c) : 0.f;
    A[i] = sc[i];
    Bc[i] = 0.f;
    D[i] = 0.f;
    if (s1) {
      const float k1 = (float)(s1[c] * inv_m), k2 = (float)(s2[c] * inv_m);
      const float mu = __ldg(mean + c), rs = __ldg(rstd + c);
      Bc[i] = -rs * k2 * sc[i];
      D[i] = (-k1 + mu * rs * k2) * sc[i];
    }
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nvec; e += stride) {
    const size_t o = (size_t)e * V;
    float dz[V], xx[V];
    vload(dy + o, dz);
    if (s1 || relu == 2) vload(x + o, xx);
    if (relu == 1) {
      float yy[V];
      vload(y + o, yy);
#pragma unroll
      for (int i = 0; i < V; ++i) if (!(yy[i] > 0.f)) dz[i] = 0.f;
    } else if (relu == 2) {
#pragma unroll
      for (int i = 0; i < V; ++i) if (!(fmaf(xx[i], sc[i], sh[i]) > 0.f)) dz[i] = 0.f;
    }
    if (dres) vstore(dres + o, dz);
    float g[V];
    if (s1) {
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = fmaf(dz[i], A[i], fmaf(xx[i], Bc[i], D[i]));
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = dz[i] * A[i];
    }
    vstore(dx + o, g);
  }
}

// colsum / period (optional, streaming kernels with period <= 32 only): also accumulate sum over rows of dx per
// (row % period, channel) into colsum[period][C] (double, zero-filled by the caller) -- the bias gradient of a layer whose
// output rows cycle through `period` joints, without another pass over dx.
__global__ void sums_to_float_kernel(int n, const double* __restrict__ src, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

// sums64 / sums32 (optional): the finished [2][C] double sums of p2r_col_bwd_stats (d beta, d gamma) are also written as
// float32 by this launch (streaming kernels) -- the parameter-dtype gradients without a conversion launch of their own.
extern "C" int p2r_bn_bwd_apply_ex(const void* dy, const void* x, const void* y, int dtype, long long M, int C,
                                   const float* mean, const float* rstd, const float* scale, const double* s1,
                                   const double* s2, int relu, void* dx, void* dres, const float* shift, double* colsum,
                                   int period, const double* sums64, float* sums32, void* stream) {
  P2R_CHECK_ARG(M >= 0 && C > 0, "p2r_bn_bwd_apply");
  const long long total = M * C;
  if (total == 0) {
    if (sums64 != nullptr && sums32 != nullptr)
      sums_to_float_kernel<<<p2r_ceil_div(2 * C, 256), 256, 0, (cudaStream_t)stream>>>(2 * C, sums64, sums32);
    return 0;
  }
  if (x != nullptr && ((relu != 1 && relu != 3) || y != nullptr) && (relu != 2 || shift != nullptr) &&
      p2r_stream_bn_ok(dtype, M, C, dy, x, relu == 3 ? nullptr : y, dx, dres))
    return p2r_stream_bn_bwd_apply(dy, x, (relu == 1 || relu == 3) ? y : nullptr, M, mean, rstd, scale, s1, s2, relu, dx, dres,
                                   shift, colsum, period, (cudaStream_t)stream, sums64, sums32);
  P2R_CHECK_ARG(relu != 3, "p2r_bn_bwd_apply (relu = 3, the bit mask, needs p2r_stream_bn_supported(dtype, M, C))");
  P2R_CHECK_ARG(colsum == nullptr, "p2r_bn_bwd_apply (fused column sums need p2r_stream_bn_supported(dtype, M, C))");
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (total + 255) / 256);
  if (dtype == 0 && vec_ok<float>(C, dy, x, y, dx, dres))
    bn_bwd_apply_vec_kernel<float><<<vec_grid(total / 4), 256, 0, (cudaStream_t)stream>>>(total / 4, C, 1.0 / (double)M, (const float*)dy, (const float*)x, (const float*)y, mean, rstd, scale, s1, s2, relu, (float*)dx, (float*)dres, shift);
  else if (dtype == 1 && vec_ok<__nv_bfloat16>(C, dy, x, y, dx, dres))
    bn_bwd_apply_vec_kernel<__nv_bfloat16><<<vec_grid(total / 8), 256, 0, (cudaStream_t)stream>>>(total / 8, C, 1.0 / (double)M, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, mean, rstd, scale, s1, s2, relu, (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, shift);
  else if (dtype == 0)
    bn_bwd_apply_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(total, C, 1.0 / (double)M, (const float*)dy, (const float*)x, (const float*)y, mean, rstd, scale, s1, s2, relu, (float*)dx, (float*)dres, shift);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(total, C, 1.0 / (double)M, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)y, mean, rstd, scale, s1, s2, relu, (__nv_bfloat16*)dx, (__nv_bfloat16*)dres, shift);
  if (sums64 != nullptr && sums32 != nullptr)
    sums_to_float_kernel<<<p2r_ceil_div(2 * C, 256), 256, 0, (cudaStream_t)stream>>>(2 * C, sums64, sums32);
  P2R_RETURN_LAUNCH("p2r_bn_bwd_apply");
}

extern "C" int p2r_bn_bwd_apply(const void* dy, const void* x, const void* y, int dtype, long long M, int C,
                                const float* mean, const float* rstd, const float* scale, const double* s1,
                                const double* s2, int relu, void* dx, void* dres, const float* shift, double* colsum,
                                int period, void* stream) {
  return p2r_bn_bwd_apply_ex(dy, x, y, dtype, M, C, mean, rstd, scale, s1, s2, relu, dx, dres, shift, colsum, period, nullptr,
                             nullptr, stream);
}

// dz = dy * (y > 0)   (backward of a ReLU fused into a GEMM epilogue)
template <typename T>
__global__ void __launch_bounds__(256)
relu_bwd_kernel(long long total, const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride)
    stf<T>(dz + e, ldf<T>(y + e) > 0.f ? ldf<T>(dy + e) : 0.f);
}

extern "C" int p2r_relu_bwd(const void* dy, const void* y, int dtype, long long total, void* dz, void* stream) {
  if (total <= 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (total + 255) / 256);
  if (dtype == 0)
    relu_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(total, (const float*)dy, (const float*)y, (float*)dz);
  else
    relu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(total, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)y, (__nv_bfloat16*)dz);
  P2R_RETURN_LAUNCH("p2r_relu_bwd");
}

// ================================================================================================
// temporal unfold / fold for the (3x1) temporal convolution (stgcn_layers.py:405-411, padding 1):
//   x [B,T,V,C] -> col [B*T*V, KT*C], col[(b,t,v), dt*C + c] = x[b, t+dt-pad, v, c] or 0 outside [0,T)
//   fold is the adjoint (gather form, no atomics): dx[b,t,v,c] = sum_dt dcol[(b,t-dt+pad,v), dt*C+c]
// ================================================================================================
template <typename T, bool FOLD>
__global__ void __launch_bounds__(256)
temporal_unfold_kernel(int Tn, int V, int C, int KT, long long total, const T* __restrict__ src, T* __restrict__ dst) {
  const int pad = (KT - 1) / 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    if (!FOLD) {
      // e indexes col: ((b*T + t)*V + v) * (KT*C) + dt*C + c
      const int c = (int)(e % C);
      long long r = e / C;
      const int dt = (int)(r % KT);
      r /= KT;  // (b*T + t)*V + v
      const int v = (int)(r % V);
      const long long bt = r / V;
      const int t = (int)(bt % Tn);
      const int ts = t + dt - pad;
      float val = 0.f;
      if (ts >= 0 && ts < Tn) val = ldf<T>(src + ((bt - t + ts) * V + v) * C + c);
      stf<T>(dst + e, val);
    } else {
      // e indexes dx: ((b*T + t)*V + v)*C + c
      const int c = (int)(e % C);
      long long r = e / C;
      const int v = (int)(r % V);
      const long long bt = r / V;
      const int t = (int)(bt % Tn);
      float acc = 0.f;
      for (int dt = 0; dt < KT; ++dt) {
        const int tr = t - dt + pad;  // row whose tap dt read x[t]
        if (tr >= 0 && tr < Tn) acc += ldf<T>(src + (((bt - t + tr) * V + v) * KT + dt) * C + c);
      }
      stf<T>(dst + e, acc);
    }
  }
}

extern "C" int p2r_temporal_unfold(const void* x, int dtype, int B, int Tn, int V, int C, int KT, void* col,
                                   void* stream) {
  P2R_CHECK_ARG(B >= 0 && Tn > 0 && V > 0 && C > 0 && KT > 0 && (KT & 1), "p2r_temporal_unfold");
  const long long total = (long long)B * Tn * V * C * KT;
  if (total == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (total + 255) / 256);
  if (dtype == 0)
    temporal_unfold_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(Tn, V, C, KT, total, (const float*)x, (float*)col);
  else
    temporal_unfold_kernel<__nv_bfloat16, false><<<grid, 256, 0, (cudaStream_t)stream>>>(Tn, V, C, KT, total, (const __nv_bfloat16*)x, (__nv_bfloat16*)col);
  P2R_RETURN_LAUNCH("p2r_temporal_unfold");
}

extern "C" int p2r_temporal_fold(const void* dcol, int dtype, int B, int Tn, int V, int C, int KT, void* dx,
                                 void* stream) {
  P2R_CHECK_ARG(B >= 0 && Tn > 0 && V > 0 && C > 0 && KT > 0 && (KT & 1), "p2r_temporal_fold");
  const long long total = (long long)B * Tn * V * C;
  if (total == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (total + 255) / 256);
  if (dtype == 0)
    temporal_unfold_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(Tn, V, C, KT, total, (const float*)dcol, (float*)dx);
  else
    temporal_unfold_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(Tn, V, C, KT, total, (const __nv_bfloat16*)dcol, (__nv_bfloat16*)dx);
  P2R_RETURN_LAUNCH("p2r_temporal_fold");
}

// ================================================================================================
// channel-last grouping for the set-abstraction layer (the layout the fused path uses instead of the
// reference's (B,C,N) -> (B,C,P,S) group_points):
//   group_rows:      feats [B,N,C], idx [B,P,S] i32 -> out [B,P,S,C]      (row gather, 128-bit copies)
//   group_rows_grad: grad  [B,P,S,C] -> dfeats [B,N,C] (fp32 atomics, zero-filled by caller)
//   maxpool_rows:    x [R,S,C] -> out [R,C], arg [R,C] u8  (first maximum wins, like F.max_pool2d)
//   maxpool_rows_grad: dout [R,C], arg -> dx [R,S,C]
// ================================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
group_rows_kernel(int N, int C, long long rows, int rows_per_batch, const T* __restrict__ feats,
                  const int* __restrict__ idx, T* __restrict__ out) {
  // one warp per output row
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  const long long b = w / rows_per_batch;
  const int src_row = __ldg(idx + w);
  const T* s = feats + ((size_t)b * N + src_row) * C;
  T* d = out + (size_t)w * C;
  constexpr int VEC = 16 / sizeof(T);
  if (C % VEC == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(s);
    uint4* d4 = reinterpret_cast<uint4*>(d);
    for (int i = lane; i < C / VEC; i += 32) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = lane; i < C; i += 32) d[i] = s[i];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
group_rows_grad_kernel(int N, int C, long long rows, int rows_per_batch, const T* __restrict__ grad,
                       const int* __restrict__ idx, float* __restrict__ dfeats) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows) return;
  const long long b = w / rows_per_batch;
  const int dst_row = __ldg(idx + w);
  float* d = dfeats + ((size_t)b * N + dst_row) * C;
  const T* g = grad + (size_t)w * C;
  for (int i = lane; i < C; i += 32) atomicAdd(d + i, ldf<T>(g + i));
}

extern "C" int p2r_group_rows(const void* feats, int dtype, const int* idx, int B, int N, int C, int P, int S, void* out,
                              void* stream) {
  P2R_CHECK_ARG(B >= 0 && N > 0 && C > 0 && P >= 0 && S >= 0, "p2r_group_rows");
  const long long rows = (long long)B * P * S;
  if (rows == 0) return 0;
  const int grid = p2r_ceil_div(rows * 32, 256);
  if (dtype == 0)
    group_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(N, C, rows, P * S, (const float*)feats, idx, (float*)out);
  else
    group_rows_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(N, C, rows, P * S, (const __nv_bfloat16*)feats, idx, (__nv_bfloat16*)out);
  P2R_RETURN_LAUNCH("p2r_group_rows");
}

// dfeats is ALWAYS float32 [B,N,C], zero-filled by the caller.
extern "C" int p2r_group_rows_grad(const void* grad, int dtype, const int* idx, int B, int N, int C, int P, int S,
                                   float* dfeats, void* stream) {
  P2R_CHECK_ARG(B >= 0 && N > 0 && C > 0 && P >= 0 && S >= 0, "p2r_group_rows_grad");
  const long long rows = (long long)B * P * S;
  if (rows == 0) return 0;
  const int grid = p2r_ceil_div(rows * 32, 256);
  if (dtype == 0)
    group_rows_grad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(N, C, rows, P * S, (const float*)grad, idx, dfeats);
  else
    group_rows_grad_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(N, C, rows, P * S, (const __nv_bfloat16*)grad, idx, dfeats);
  P2R_RETURN_LAUNCH("p2r_group_rows_grad");
}

// Adjoint of out[b][p] = feats[b][idx[b][p]] (one picked row per slot: group_rows with S = 1) written destination-major:
// one warp per row of dfeats scans the P picks of its batch, sums the matching gradient rows (a frame picked twice
// gets both) and stores the row -- zeros included, so the caller does not zero-fill and nothing is atomic.
template <typename T>
__global__ void __launch_bounds__(256)
select_rows_grad_kernel(int N, int C, int P, long long dst_rows, const T* __restrict__ grad,
                        const int* __restrict__ idx, T* __restrict__ dfeats) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= dst_rows) return;
  const long long b = w / N;
  const int row = (int)(w - b * N);
  const int* picks = idx + b * P;
  constexpr int VEC = 16 / sizeof(T);
  T* d = dfeats + (size_t)w * C;
  __shared__ int s_match[8][32];                 // the first 32 slots that picked this row, per warp
  int* match = s_match[threadIdx.x >> 5];
  int count = 0;
  for (int pc = 0; pc < P; pc += 512) {          // 16 independent loads per lane, then their ballots: one round trip per
    int pk[16];                                  // 512 picks instead of sixteen dependent ones
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int p = pc + 32 * k + lane;
      pk[k] = p < P ? __ldg(picks + p) : -1;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const bool hit = pk[k] == row;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m != 0u) {
        if (hit) {
          const int slot = count + __popc(m & ((1u << lane) - 1u));
          if (slot < 32) match[slot] = pc + 32 * k + lane;
        }
        count += __popc(m);
      }
    }
  }
  __syncwarp();
  if (count == 0) {
    if (C % VEC == 0) {
      uint4* d4 = reinterpret_cast<uint4*>(d);
      for (int i = lane; i < C / VEC; i += 32) d4[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int i = lane; i < C; i += 32) stf<T>(d + i, 0.f);
    }
  } else if (count == 1 && C % VEC == 0) {
    const uint4* g4 = reinterpret_cast<const uint4*>(grad + ((size_t)b * P + match[0]) * C);
    uint4* d4 = reinterpret_cast<uint4*>(d);
    const int nv = C / VEC;
    for (int i0 = 0; i0 < nv; i0 += 8 * 32) {    // up to eight 16-byte loads in flight per lane, then the stores
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + 32 * k + lane;
        if (i < nv) v[k] = __ldg(g4 + i);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + 32 * k + lane;
        if (i < nv) d4[i] = v[k];
      }
    }
  } else if (count <= 32 && C % VEC == 0 && sizeof(T) == 2) {
    // a frame picked several times (a standing person: arc-length sampling repeats the frame): 16-byte loads, the picks
    // of one chunk are independent loads, summed in slot order (deterministic)
    for (int i = lane; i < C / VEC; i += 32) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int q = 0; q < count; ++q) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(grad + ((size_t)b * P + match[q]) * C) + i);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[2 * j] += __uint_as_float(u[j] << 16);
          acc[2 * j + 1] += __uint_as_float(u[j] & 0xffff0000u);
        }
      }
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        o[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      reinterpret_cast<uint4*>(d)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else if (count <= 32) {
    for (int i = lane; i < C; i += 32) {
      float t = 0.f;
      for (int q = 0; q < count; ++q) t += ldf<T>(grad + ((size_t)b * P + match[q]) * C + i);   // slot order: deterministic
      stf<T>(d + i, t);
    }
  } else {                                       // a frame picked more than 32 times: scan again per element
    for (int i = lane; i < C; i += 32) {
      float t = 0.f;
      for (int p = match[0]; p < P; ++p)
        if (__ldg(picks + p) == row) t += ldf<T>(grad + ((size_t)b * P + p) * C + i);
      stf<T>(d + i, t);
    }
  }
}

extern "C" int p2r_select_rows_grad(const void* grad, int dtype, const int* idx, int B, int N, int C, int P, void* dfeats,
                                    void* stream) {
  P2R_CHECK_ARG(B >= 0 && N > 0 && C > 0 && P >= 0 && (dtype == 0 || dtype == 1), "p2r_select_rows_grad");
  const long long rows = (long long)B * N;
  if (rows == 0) return 0;
  const int grid = p2r_ceil_div(rows * 32, 256);
  if (dtype == 0)
    select_rows_grad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(N, C, P, rows, (const float*)grad, idx, (float*)dfeats);
  else
    select_rows_grad_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(N, C, P, rows, (const __nv_bfloat16*)grad, idx, (__nv_bfloat16*)dfeats);
  P2R_RETURN_LAUNCH("p2r_select_rows_grad");
}

template <typename T, bool GRAD>
__global__ void __launch_bounds__(256)
maxpool_rows_kernel(long long R, int S, int C, const T* __restrict__ src, unsigned char* __restrict__ arg,
                    T* __restrict__ dst) {
  const long long total = R * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long r = e / C;
    const int c = (int)(e % C);
    if (!GRAD) {
      const T* p = src + (size_t)r * S * C + c;
      float best = ldf<T>(p);
      int bi = 0;
      for (int s = 1; s < S; ++s) {
        const float v = ldf<T>(p + (size_t)s * C);
        if (v > best) { best = v; bi = s; }
      }
      stf<T>(dst + e, best);
      arg[e] = (unsigned char)bi;
    } else {
      const float g = ldf<T>(src + e);
      const int a = arg[e];
      T* p = dst + (size_t)r * S * C + c;
      for (int s = 0; s < S; ++s) stf<T>(p + (size_t)s * C, s == a ? g : 0.f);
    }
  }
}

extern "C" int p2r_maxpool_rows(const void* x, int dtype, long long R, int S, int C, void* out, unsigned char* arg,
                                void* stream) {
  P2R_CHECK_ARG(R >= 0 && S > 0 && S <= 255 && C > 0, "p2r_maxpool_rows");
  if (R == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (R * C + 255) / 256);
  if (dtype == 0)
    maxpool_rows_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(R, S, C, (const float*)x, arg, (float*)out);
  else
    maxpool_rows_kernel<__nv_bfloat16, false><<<grid, 256, 0, (cudaStream_t)stream>>>(R, S, C, (const __nv_bfloat16*)x, arg, (__nv_bfloat16*)out);
  P2R_RETURN_LAUNCH("p2r_maxpool_rows");
}

extern "C" int p2r_maxpool_rows_grad(const void* dout, int dtype, const unsigned char* arg, long long R, int S, int C,
                                     void* dx, void* stream) {
  P2R_CHECK_ARG(R >= 0 && S > 0 && S <= 255 && C > 0, "p2r_maxpool_rows_grad");
  if (R == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (R * C + 255) / 256);
  if (dtype == 0)
    maxpool_rows_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(R, S, C, (const float*)dout, const_cast<unsigned char*>(arg), (float*)dx);
  else
    maxpool_rows_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(R, S, C, (const __nv_bfloat16*)dout, const_cast<unsigned char*>(arg), (__nv_bfloat16*)dx);
  P2R_RETURN_LAUNCH("p2r_maxpool_rows_grad");
}


// ================================================================================================
// embed_sum: x[f,j,:] = sk[f,j,:] + mean_k pos[f,k,:]   (models/p2rnet/modules/stgcn.py:121,129: the relative-position
// embedding averaged over the 20-frame window is broadcast onto every joint feature of the frame).
// One warp per frame f = (b,t); fp32 accumulation, one rounding.  Backward: dsk = dx (aliased by the caller),
// dpos[f,k,:] = (1/K) sum_j dx[f,j,:].
// ================================================================================================
template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
embed_sum_kernel(long long frames, int J, int K, int C, const T* __restrict__ a, const T* __restrict__ b_in,
                 T* __restrict__ out) {
  const long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= frames) return;
  const float inv_k = 1.f / (float)K;
  for (int c = lane; c < C; c += 32) {
    if (!BWD) {
      // a = sk [frames,J,C], b_in = pos [frames,K,C], out = x [frames,J,C]
      float m = 0.f;
      for (int k = 0; k < K; ++k) m += ldf<T>(b_in + ((size_t)f * K + k) * C + c);
      m *= inv_k;
      for (int j = 0; j < J; ++j) {
        const size_t o = ((size_t)f * J + j) * C + c;
        stf<T>(out + o, ldf<T>(a + o) + m);
      }
    } else {
      // a = dx [frames,J,C], out = dpos [frames,K,C]
      float sum = 0.f;
      for (int j = 0; j < J; ++j) sum += ldf<T>(a + ((size_t)f * J + j) * C + c);
      sum *= inv_k;
      for (int k = 0; k < K; ++k) stf<T>(out + ((size_t)f * K + k) * C + c, sum);
    }
  }
}

// C = 64 bf16: 16-byte vectors.  A warp owns one frame; lane = (row lane 0..3, 8-channel group 0..7): every load / store
// instruction of the warp moves four whole 128-byte rows.
template <bool BWD>
__global__ void __launch_bounds__(256)
embed_sum_vec64_kernel(long long frames, int J, int K, const __nv_bfloat16* __restrict__ a,
                       const __nv_bfloat16* __restrict__ b_in, __nv_bfloat16* __restrict__ out) {
  const long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= frames) return;
  const int rl = lane >> 3, c0 = (lane & 7) * 8;
  const int nsrc = BWD ? J : K, ndst = BWD ? K : J;
  const __nv_bfloat16* src = (BWD ? a : b_in) + (size_t)f * nsrc * 64 + c0;
  float m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = 0.f;
  for (int r = rl; r < nsrc; r += 4) {
    float v[8];
    vload(src + (size_t)r * 64, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    m[i] += __shfl_xor_sync(0xffffffffu, m[i], 8);
    m[i] += __shfl_xor_sync(0xffffffffu, m[i], 16);
    m[i] *= 1.f / (float)K;
  }
  __nv_bfloat16* dst = out + (size_t)f * ndst * 64 + c0;
  if (BWD) {
    for (int r = rl; r < ndst; r += 4) vstore(dst + (size_t)r * 64, m);
  } else {
    const __nv_bfloat16* skp = a + (size_t)f * J * 64 + c0;
    for (int r = rl; r < ndst; r += 4) {
      float v[8];
      vload(skp + (size_t)r * 64, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += m[i];
      vstore(dst + (size_t)r * 64, v);
    }
  }
}

static bool embed_vec_ok(int dtype, int C, const void* p0, const void* p1, const void* p2) {
  auto al = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return dtype == 1 && C == 64 && al(p0) && al(p1) && al(p2);
}

extern "C" int p2r_embed_sum(const void* sk, const void* pos, int dtype, long long frames, int J, int K, int C, void* x,
                             void* stream) {
  P2R_CHECK_ARG(frames >= 0 && J > 0 && K > 0 && C > 0, "p2r_embed_sum");
  if (frames == 0) return 0;
  const int grid = p2r_ceil_div(frames * 32, 256);
  if (embed_vec_ok(dtype, C, sk, pos, x)) {
    embed_sum_vec64_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, J, K, (const __nv_bfloat16*)sk, (const __nv_bfloat16*)pos, (__nv_bfloat16*)x);
    P2R_RETURN_LAUNCH("p2r_embed_sum");
  }
  if (dtype == 0)
    embed_sum_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, J, K, C, (const float*)sk, (const float*)pos, (float*)x);
  else
    embed_sum_kernel<__nv_bfloat16, false><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, J, K, C, (const __nv_bfloat16*)sk, (const __nv_bfloat16*)pos, (__nv_bfloat16*)x);
  P2R_RETURN_LAUNCH("p2r_embed_sum");
}

extern "C" int p2r_embed_sum_grad(const void* dx, int dtype, long long frames, int J, int K, int C, void* dpos,
                                  void* stream) {
  P2R_CHECK_ARG(frames >= 0 && J > 0 && K > 0 && C > 0, "p2r_embed_sum_grad");
  if (frames == 0) return 0;
  const int grid = p2r_ceil_div(frames * 32, 256);
  if (embed_vec_ok(dtype, C, dx, dpos, nullptr)) {
    embed_sum_vec64_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, J, K, (const __nv_bfloat16*)dx, nullptr, (__nv_bfloat16*)dpos);
    P2R_RETURN_LAUNCH("p2r_embed_sum_grad");
  }
  if (dtype == 0)
    embed_sum_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, J, K, C, (const float*)dx, nullptr, (float*)dpos);
  else
    embed_sum_kernel<__nv_bfloat16, true><<<grid, 256, 0, (cudaStream_t)stream>>>(frames, J, K, C, (const __nv_bfloat16*)dx, nullptr, (__nv_bfloat16*)dpos);
  P2R_RETURN_LAUNCH("p2r_embed_sum_grad");
}


// ================================================================================================
// small-K linear layers (the 3 -> 64 first layers of pos_embed / sk_feat, stgcn.py:45-50): K <= 4 input channels,
// so the layer is a memory-bound elementwise map, not a GEMM.
//   forward : y[m, n] = sum_c x[m, c] * W[n, c] (+ bias[n])        thread = VEC outputs of one row, W in registers
//   d weight: dW[n, c] = sum_m dz[m, n] * x[m, c]                  column reduction with K accumulators per channel
// (no input gradient: the inputs are data).  N % VEC == 0 and (256 * VEC) % N == 0.
// ================================================================================================
template <typename T, typename TX = T>
__global__ void __launch_bounds__(256)
smallk_linear_kernel(long long M, int N, int K, const TX* __restrict__ x, const float* __restrict__ W,
                     const float* __restrict__ bias, T* __restrict__ y) {
  constexpr int V = VecN<T>::N;
  const int n0 = (threadIdx.x * V) % N;
  const int tpr = N / V;
  float w[V][4], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    b[i] = bias ? __ldg(bias + n0 + i) : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) w[i][c] = c < K ? __ldg(W + (size_t)(n0 + i) * K + c) : 0.f;
  }
  const long long nvec = M * tpr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nvec; e += stride) {
    const long long m = e / tpr;
    float xv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < K; ++c) xv[c] = ldf<TX>(x + (size_t)m * K + c);
    float o[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float t = b[i];
#pragma unroll
      for (int c = 0; c < 4; ++c) t = fmaf(xv[c], w[i][c], t);
      o[i] = t;
    }
    vstore(y + (size_t)m * N + n0, o);
  }
}

template <typename T, typename TX = T>
__global__ void __launch_bounds__(256)
smallk_dw_kernel(long long M, int N, int K, const T* __restrict__ dz, const TX* __restrict__ x, int rows_per_cta,
                 float* __restrict__ dW) {
  constexpr int V = VecN<T>::N;
  __shared__ float sh[256 * V];
  const int tpr = N / V, rl = 256 / tpr;
  const int cl = threadIdx.x % tpr, r_lane = threadIdx.x / tpr;
  const int n0 = cl * V;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float acc[4][V];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[c][i] = 0.f;
  for (long long r = r0 + r_lane; r < r1; r += rl) {
    float g[V];
    vload(dz + (size_t)r * N + n0, g);
    float xv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < K; ++c) xv[c] = ldf<TX>(x + (size_t)r * K + c);
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[c][i] = fmaf(g[i], xv[c], acc[c][i]);
  }
  for (int c = 0; c < K; ++c) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) sh[threadIdx.x * V + i] = acc[c][i];
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += 256) {
      const int t = n / V, i = n % V;
      float tot = 0.f;
      for (int l = 0; l < rl; ++l) tot += sh[(l * tpr + t) * V + i];
      atomicAdd(dW + (size_t)n * K + c, tot);
    }
  }
}

extern "C" int p2r_smallk_linear(const void* x, const float* W, const float* bias, int dtype, long long M, int N, int K,
                                 void* y, void* stream) {
  P2R_CHECK_ARG(M >= 0 && K >= 1 && K <= 4 && N > 0, "p2r_smallk_linear");
  const int V = dtype == 0 ? 4 : 8;
  P2R_CHECK_ARG(N % V == 0 && (256 * V) % N == 0, "p2r_smallk_linear (N must divide 256*VEC)");
  if (M == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (M * (N / V) + 255) / 256);
  if (dtype == 0)
    smallk_linear_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, (const float*)x, W, bias, (float*)y);
  else
    smallk_linear_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, (const __nv_bfloat16*)x, W, bias, (__nv_bfloat16*)y);
  P2R_RETURN_LAUNCH("p2r_smallk_linear");
}

// The first layer of a point MLP in throughput mode: COORDINATES stay float32 (rounding a position of ~1 m to bf16 moves it
// by millimetres: the largest single contribution to the bf16-vs-fp32 error of the whole backbone), the output is bf16.
extern "C" int p2r_smallk_linear_mixed(const float* x, const float* W, const float* bias, long long M, int N, int K,
                                       void* y, void* stream) {
  P2R_CHECK_ARG(M >= 0 && K >= 1 && K <= 4 && N > 0, "p2r_smallk_linear_mixed");
  P2R_CHECK_ARG(N % 8 == 0 && (256 * 8) % N == 0, "p2r_smallk_linear_mixed (N must divide 2048)");
  if (M == 0) return 0;
  const int grid = (int)min((long long)P2R_SM_COUNT * 16, (M * (N / 8) + 255) / 256);
  smallk_linear_kernel<__nv_bfloat16, float><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, x, W, bias, (__nv_bfloat16*)y);
  P2R_RETURN_LAUNCH("p2r_smallk_linear_mixed");
}

extern "C" int p2r_smallk_dw_mixed(const void* dz, const float* x, long long M, int N, int K, float* dW, void* stream) {
  P2R_CHECK_ARG(M >= 0 && K >= 1 && K <= 4 && N > 0, "p2r_smallk_dw_mixed");
  P2R_CHECK_ARG(N % 8 == 0 && 256 % (N / 8) == 0, "p2r_smallk_dw_mixed");
  if (M == 0) return 0;
  int rpc;
  const int grid = colreduce_grid(M, &rpc);
  smallk_dw_kernel<__nv_bfloat16, float><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, (const __nv_bfloat16*)dz, x, rpc, dW);
  P2R_RETURN_LAUNCH("p2r_smallk_dw_mixed");
}

// dW f32[N,K] must be zero-filled by the caller.
extern "C" int p2r_smallk_dw(const void* dz, const void* x, int dtype, long long M, int N, int K, float* dW,
                             void* stream) {
  P2R_CHECK_ARG(M >= 0 && K >= 1 && K <= 4 && N > 0, "p2r_smallk_dw");
  const int V = dtype == 0 ? 4 : 8;
  P2R_CHECK_ARG(N % V == 0 && 256 % (N / V) == 0, "p2r_smallk_dw");
  if (M == 0) return 0;
  int rpc;
  const int grid = colreduce_grid(M, &rpc);
  if (dtype == 0)
    smallk_dw_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, (const float*)dz, (const float*)x, rpc, dW);
  else
    smallk_dw_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, (const __nv_bfloat16*)dz, (const __nv_bfloat16*)x, rpc, dW);
  P2R_RETURN_LAUNCH("p2r_smallk_dw");
}


// column sums of a WIDE matrix [M, C] (C a multiple of VEC, any size: the 1600-wide graph-conv output whose bias
// gradient is needed): grid.x tiles the columns (32 vectors per CTA), grid.y tiles the rows; s1 zero-filled by the caller.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_wide_kernel(long long M, int C, const T* __restrict__ dy, int rows_per_cta, double* __restrict__ s1) {
  constexpr int V = VecN<T>::N;
  __shared__ float sh[256 * V];
  const int cv = threadIdx.x & 31, r_lane = threadIdx.x >> 5;   // 32 column vectors x 8 row lanes
  const int c0 = (blockIdx.x * 32 + cv) * V;
  const long long r0 = (long long)blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float acc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc[i] = 0.f;
  if (c0 < C) {
#pragma unroll 4
    for (long long r = r0 + r_lane; r < r1; r += 8) {
      float v[V];
      vload(dy + (size_t)r * C + c0, v);
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sh[threadIdx.x * V + i] = acc[i];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * V; e += 256) {
    const int t = e / V, i = e % V;
    const int c = (blockIdx.x * 32 + t) * V + i;
    if (c < C) {
      double tot = 0.0;
      for (int l = 0; l < 8; ++l) tot += (double)sh[(l * 32 + t) * V + i];
      atomicAdd(s1 + c, tot);
    }
  }
}

extern "C" int p2r_col_sum_wide(const void* dy, int dtype, long long M, int C, double* s1, void* stream) {
  const int V = dtype == 0 ? 4 : 8;
  P2R_CHECK_ARG(M >= 0 && C > 0 && C % V == 0, "p2r_col_sum_wide");
  if (M == 0) return 0;
  // [M, V * 64] bf16 (the graph convolution's output gradient): rows of 64 channels that cycle through V joints -> the
  // bulk-TMA ring with register-resident sums (stream_bn.cu)
  if (C % 64 == 0 && C / 64 >= 2 && C / 64 <= 32 && p2r_stream_bn_ok(dtype, M * (C / 64), 64, dy))
    return p2r_stream_colsum_period(dy, M * (C / 64), C / 64, s1, (cudaStream_t)stream);
  const int col_tiles = p2r_ceil_div(C, 32 * V);
  int row_tiles = (P2R_SM_COUNT * 8 + col_tiles - 1) / col_tiles;
  long long rpc = (M + row_tiles - 1) / row_tiles;
  if (rpc < 64) rpc = 64;
  row_tiles = (int)((M + rpc - 1) / rpc);
  dim3 grid(col_tiles, row_tiles);
  if (dtype == 0)
    colsum_wide_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const float*)dy, (int)rpc, s1);
  else
    colsum_wide_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, (const __nv_bfloat16*)dy, (int)rpc, s1);
  P2R_RETURN_LAUNCH("p2r_col_sum_wide");
}
