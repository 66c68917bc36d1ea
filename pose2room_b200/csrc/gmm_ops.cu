// gmm_ops.cu -- training-time point prediction of a Gaussian-mixture box head and its gradient, one launch each.
//
// ref: models/p2rnet/modules/mdn.py:31-84 (MixtureDensityHead.forward / sample / generate_samples / point_prediction with
// n_samples = 1): out[r,:] = sum_g sigmoid(logit[r,g]) * (mu[g,:] + exp(log_sigma[g,:]) * eps[r,g,:]).  The reference (and
// the torch path in pose2room_b200/p2rnet/mdn.py) runs this as ~8 elementwise / reduction kernels forward and ~20 backward
// on (rows, G, 1, d) tensors, three heads per step; here it is one kernel per direction.  eps is still drawn by torch with
// the reference's call (same RNG consumption).  Arithmetic: gmm_math.h (shared with the host-compiled CPU test).
//
//   forward : one warp per row, lanes stride over the G components, warp-shuffle sum.
//   backward: one thread per component g, a block walks `rows_per_block` rows: d logit[r,g] is written coalesced, d mu /
//             d log_sigma accumulate in registers; per-block partials -> global, the last block to finish adds them in
//             block order (deterministic) and writes d mu / d log_sigma.
// Opt-in (P2R_FUSED_GMM=1) until it has been A/B-ed on a B200.
#include "p2r_common.cuh"
#include "p2r_b200.h"
#include "gmm_math.h"

template <typename T>
__device__ __forceinline__ float p2rg_load_logit(const T* p);
template <>
__device__ __forceinline__ float p2rg_load_logit<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float p2rg_load_logit<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void p2rg_store_logit(T* p, double v);
template <>
__device__ __forceinline__ void p2rg_store_logit<float>(float* p, double v) { *p = (float)v; }
template <>
__device__ __forceinline__ void p2rg_store_logit<__nv_bfloat16>(__nv_bfloat16* p, double v) { *p = __float2bfloat16_rn((float)v); }

// mu (MuT) and sigma = exp(log_sigma) (float32 like torch.exp on the float32 parameter) as float64 in shared memory
template <typename MuT>
__device__ __forceinline__ void p2rg_stage_params(const MuT* mu, const float* log_sigma, int n, double* s_mu, double* s_sig) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_mu[i] = (double)mu[i];
    s_sig[i] = (double)expf(log_sigma[i]);
  }
  __syncthreads();
}

template <typename MuT, typename LogitT>
__global__ void __launch_bounds__(256)
gmm_mix_fwd_kernel(const LogitT* __restrict__ logits, const MuT* __restrict__ mu, const float* __restrict__ log_sigma,
                   const MuT* __restrict__ eps, long long rows, int G, int D, MuT* __restrict__ out) {
  P2R_DYN_SMEM(double, s_par);                          // mu[G*D] | sigma[G*D]
  double* s_mu = s_par;
  double* s_sig = s_par + G * D;
  p2rg_stage_params(mu, log_sigma, G * D, s_mu, s_sig);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * wpb + warp; r < rows; r += (long long)gridDim.x * wpb) {
    double acc[P2RG_MAX_D] = {0.0, 0.0, 0.0, 0.0};
    for (int g = lane; g < G; g += 32) {
      const double pi = p2rg_sigmoid(p2rg_load_logit(logits + r * G + g));
      double e[P2RG_MAX_D];
      for (int c = 0; c < D; ++c) e[c] = (double)eps[(r * G + g) * D + c];
      p2rg_accumulate(pi, s_mu + g * D, s_sig + g * D, e, D, acc);
    }
    for (int c = 0; c < D; ++c) {
      double v = acc[c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) out[r * D + c] = (MuT)v;
    }
  }
}

#define P2RG_ROW_GROUPS 4
template <typename MuT, typename LogitT>
__global__ void __launch_bounds__(1024)
gmm_mix_bwd_kernel(const LogitT* __restrict__ logits, const MuT* __restrict__ mu, const float* __restrict__ log_sigma,
                   const MuT* __restrict__ eps, const MuT* __restrict__ dout, long long rows, int G, int D,
                   int rows_per_block, LogitT* __restrict__ dlogits, double* __restrict__ partials,
                   unsigned int* __restrict__ counter, MuT* __restrict__ dmu, float* __restrict__ dls) {
  // Round 2: P2RG_ROW_GROUPS x as many threads per block (thread = (component g, row group q); a row group walks
  // rows_per_block / P2RG_ROW_GROUPS rows), the groups' partial sums combined through shared memory in a fixed order --
  // the one-thread-per-component version walked 32 rows per thread with 128 mostly idle blocks: ~200 us inside a step.
  P2R_DYN_SMEM(double, s_par);                // mu[G*D] | sigma[G*D] | dout[rows_per_block*D] | acc[ROW_GROUPS][G][2*D]
  double* s_mu = s_par;
  double* s_sig = s_par + G * D;
  double* s_dout = s_sig + G * D;
  double* s_acc = s_dout + rows_per_block * D;
  __shared__ bool s_last;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const int nr = (int)((rows - r0) < rows_per_block ? (rows - r0) : rows_per_block);
  for (int i = threadIdx.x; i < nr * D; i += blockDim.x) s_dout[i] = (double)dout[r0 * D + i];
  p2rg_stage_params(mu, log_sigma, G * D, s_mu, s_sig);       // ends with __syncthreads
  const int gpad = blockDim.x / P2RG_ROW_GROUPS;
  const int g = threadIdx.x % gpad, q = threadIdx.x / gpad;
  const int per_q = rows_per_block / P2RG_ROW_GROUPS;
  if (g < G) {
    double am[P2RG_MAX_D] = {0.0, 0.0, 0.0, 0.0}, al[P2RG_MAX_D] = {0.0, 0.0, 0.0, 0.0};
    const int i1 = (q + 1) * per_q < nr ? (q + 1) * per_q : nr;
#pragma unroll 4
    for (int i = q * per_q; i < i1; ++i) {
      const long long r = r0 + i;
      const double pi = p2rg_sigmoid(p2rg_load_logit(logits + r * G + g));
      double e[P2RG_MAX_D];
      for (int c = 0; c < D; ++c) e[c] = (double)eps[(r * G + g) * D + c];
      p2rg_store_logit(dlogits + r * G + g, p2rg_backward(pi, s_mu + g * D, s_sig + g * D, e, s_dout + i * D, D, am, al));
    }
    double* pa = s_acc + ((size_t)q * G + g) * 2 * D;
    for (int c = 0; c < D; ++c) {
      pa[c] = am[c];
      pa[D + c] = al[c];
    }
  }
  __syncthreads();
  if (q == 0 && g < G) {
    double* pm = partials + (size_t)blockIdx.x * 2 * G * D;
    for (int c = 0; c < D; ++c) {
      double tm = 0.0, tl = 0.0;
#pragma unroll
      for (int qq = 0; qq < P2RG_ROW_GROUPS; ++qq) {           // fixed order: deterministic
        tm += s_acc[((size_t)qq * G + g) * 2 * D + c];
        tl += s_acc[((size_t)qq * G + g) * 2 * D + D + c];
      }
      pm[g * D + c] = tm;
      pm[G * D + g * D + c] = tl;
    }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // deterministic tail: eight independent partial sums per output (eight L2 loads in flight instead of a chain of
  // gridDim.x dependent ones: this loop was ~80 of the kernel's 100 us), combined in a fixed order
  const size_t pstride = (size_t)2 * G * D;
  for (int i = threadIdx.x; i < 2 * G * D; i += blockDim.x) {
    double t[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    unsigned blk = 0;
    for (; blk + 8 <= gridDim.x; blk += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] += __ldcg(partials + (size_t)(blk + u) * pstride + i);
    }
    for (; blk < gridDim.x; ++blk) t[0] += __ldcg(partials + (size_t)blk * pstride + i);
    const double tot = ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
    if (i < G * D) dmu[i] = (MuT)tot;
    else dls[i - G * D] = (float)tot;
  }
}

#define P2RG_ROWS_PER_BLOCK 32

extern "C" long long p2r_gmm_mix_workspace(long long rows, int g, int d) {
  const long long blocks = (rows + P2RG_ROWS_PER_BLOCK - 1) / P2RG_ROWS_PER_BLOCK;
  return 1 + blocks * 2 * g * d;
}

static int gmm_check(long long rows, int g, int d, const char* where) {
  if (!(rows >= 0 && g > 0 && g <= P2RG_MAX_G && d > 0 && d <= P2RG_MAX_D)) {
    p2r_set_last_error(where, -1);
    return -1;
  }
  return 0;
}

extern "C" int p2r_gmm_mix(const void* logits, int logits_bf16, const void* mu, int mu_f64, const float* log_sigma,
                           const void* eps, long long rows, int g, int d, void* out, void* stream) {
  if (gmm_check(rows, g, d, "p2r_gmm_mix: bad argument")) return -1;
  if (rows == 0) return 0;
  const size_t smem = (size_t)2 * g * d * sizeof(double);
  int grid = p2r_ceil_div(rows, 8);
  if (grid > P2R_SM_COUNT * 8) grid = P2R_SM_COUNT * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define P2RG_FWD(MuT, LogitT)                                                                                       \
  do {                                                                                                              \
    auto kern = gmm_mix_fwd_kernel<MuT, LogitT>;                                                                    \
    P2R_LAUNCH(kern, grid, 256, smem, st, (const LogitT*)logits, (const MuT*)mu, log_sigma, (const MuT*)eps, rows,  \
               g, d, (MuT*)out);                                                                                    \
  } while (0)
  if (mu_f64) { if (logits_bf16) P2RG_FWD(double, __nv_bfloat16); else P2RG_FWD(double, float); }
  else { if (logits_bf16) P2RG_FWD(float, __nv_bfloat16); else P2RG_FWD(float, float); }
#undef P2RG_FWD
  P2R_RETURN_LAUNCH("p2r_gmm_mix");
}

extern "C" int p2r_gmm_mix_grad(const void* logits, int logits_bf16, const void* mu, int mu_f64, const float* log_sigma,
                                const void* eps, const void* dout, long long rows, int g, int d, void* dlogits,
                                void* dmu, float* dls, double* workspace, long long workspace_doubles, void* stream) {
  if (gmm_check(rows, g, d, "p2r_gmm_mix_grad: bad argument")) return -1;
  P2R_CHECK_ARG(rows > 0, "p2r_gmm_mix_grad");
  P2R_CHECK_ARG(workspace_doubles >= p2r_gmm_mix_workspace(rows, g, d), "p2r_gmm_mix_grad (workspace too small)");
  const int grid = p2r_ceil_div(rows, P2RG_ROWS_PER_BLOCK);
  const int threads = P2RG_ROW_GROUPS * ((g + 31) / 32 * 32);        // <= 4 * 256 = 1024
  const size_t smem = (size_t)(2 * g * d + P2RG_ROWS_PER_BLOCK * d + P2RG_ROW_GROUPS * g * 2 * d) * sizeof(double);
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);     // first 8 bytes; zero on entry
  double* partials = workspace + 1;
  cudaStream_t st = (cudaStream_t)stream;
#define P2RG_BWD(MuT, LogitT)                                                                                       \
  do {                                                                                                              \
    auto kern = gmm_mix_bwd_kernel<MuT, LogitT>;                                                                    \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    P2R_LAUNCH(kern, grid, threads, smem, st, (const LogitT*)logits, (const MuT*)mu, log_sigma, (const MuT*)eps,    \
               (const MuT*)dout, rows, g, d, P2RG_ROWS_PER_BLOCK, (LogitT*)dlogits, partials, counter, (MuT*)dmu,   \
               dls);                                                                                                \
  } while (0)
  if (mu_f64) { if (logits_bf16) P2RG_BWD(double, __nv_bfloat16); else P2RG_BWD(double, float); }
  else { if (logits_bf16) P2RG_BWD(float, __nv_bfloat16); else P2RG_BWD(float, float); }
#undef P2RG_BWD
  P2R_RETURN_LAUNCH("p2r_gmm_mix_grad");
}
