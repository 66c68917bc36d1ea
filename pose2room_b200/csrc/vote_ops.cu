// vote_ops.cu -- the tail of the voting module as one launch per direction.
//
// ref: models/p2rnet/modules/vote_center.py:52-58 (vote_xyz = seed_xyz + net[..., 0:3]; vote_features = seed_features +
// net[..., 3:]) and models/p2rnet/modules/network.py:89-90 (features / ||features||_2).  As torch ops that is 5 launches
// forward and ~12 backward over (B*S, 259) / (B*S, 256) tensors, all on the step's critical path; here: one warp per
// seed row, 16-byte-free simple loads (rows of 259 values are not vector-aligned), two warp-shuffle reductions.
// Arithmetic: vote_math.h.  Opt-in (P2R_FUSED_VOTE=1) until it has been A/B-ed on a B200.
#include "p2r_common.cuh"
#include "p2r_b200.h"
#include "vote_math.h"

template <typename T>
__device__ __forceinline__ float p2rv_load(const T* p);
template <>
__device__ __forceinline__ float p2rv_load<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float p2rv_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void p2rv_store(T* p, float v);
template <>
__device__ __forceinline__ void p2rv_store<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void p2rv_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float p2rv_warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return __shfl_sync(0xffffffffu, v, 0);
}

// net [R, 3+C] (NetT), seed_xyz rows `xyz_stride` floats apart, seed_feat [R, C] f32 -> vote_xyz [R,3], vote_feat [R,C],
// norm [R] (saved for the backward)
template <typename NetT>
__global__ void __launch_bounds__(256)
vote_tail_fwd_kernel(const NetT* __restrict__ net, const float* __restrict__ seed_xyz, long long xyz_stride,
                     const float* __restrict__ seed_feat, long long rows, int C, float* __restrict__ vote_xyz,
                     float* __restrict__ vote_feat, float* __restrict__ norm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * wpb + warp; r < rows; r += (long long)gridDim.x * wpb) {
    const NetT* nr = net + r * (3 + C);
    if (lane < 3) vote_xyz[r * 3 + lane] = seed_xyz[r * xyz_stride + lane] + p2rv_load(nr + lane);
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = seed_feat[r * C + c] + p2rv_load(nr + 3 + c);
      ss += v * v;
    }
    const float n = sqrtf(p2rv_warp_sum(ss));
    for (int c = lane; c < C; c += 32) {
      const float v = seed_feat[r * C + c] + p2rv_load(nr + 3 + c);
      vote_feat[r * C + c] = p2rv_normalise(v, n);
    }
    if (lane == 0) norm[r] = n;
  }
}

// g_xyz [R,3], g_feat [R,C] upstream; vote_feat / norm from the forward -> d_net [R,3+C] (NetT), d_seed_feat [R,C] f32
template <typename NetT>
__global__ void __launch_bounds__(256)
vote_tail_bwd_kernel(const float* __restrict__ g_xyz, const float* __restrict__ g_feat, const float* __restrict__ vote_feat,
                     const float* __restrict__ norm, long long rows, int C, NetT* __restrict__ d_net,
                     float* __restrict__ d_seed_feat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * wpb + warp; r < rows; r += (long long)gridDim.x * wpb) {
    NetT* dr = d_net + r * (3 + C);
    if (lane < 3) p2rv_store(dr + lane, g_xyz ? g_xyz[r * 3 + lane] : 0.f);
    float dot = 0.f;
    if (g_feat)
      for (int c = lane; c < C; c += 32) dot += vote_feat[r * C + c] * g_feat[r * C + c];
    dot = p2rv_warp_sum(dot);
    const float n = norm[r];
    for (int c = lane; c < C; c += 32) {
      const float dv = g_feat ? p2rv_dnormalise(g_feat[r * C + c], vote_feat[r * C + c], dot, n) : 0.f;
      p2rv_store(dr + 3 + c, dv);
      d_seed_feat[r * C + c] = dv;
    }
  }
}

static int vote_grid(long long rows) {
  int grid = p2r_ceil_div(rows, 8);
  return grid > P2R_SM_COUNT * 16 ? P2R_SM_COUNT * 16 : grid;
}

extern "C" int p2r_vote_tail(const void* net, int net_bf16, const float* seed_xyz, long long xyz_stride,
                             const float* seed_feat, long long rows, int c, float* vote_xyz, float* vote_feat,
                             float* norm, void* stream) {
  P2R_CHECK_ARG(rows >= 0 && c > 0 && xyz_stride >= 3, "p2r_vote_tail");
  if (rows == 0) return 0;
  const int grid = vote_grid(rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (net_bf16) {
    auto kern = vote_tail_fwd_kernel<__nv_bfloat16>;
    P2R_LAUNCH(kern, grid, 256, 0, st, (const __nv_bfloat16*)net, seed_xyz, xyz_stride, seed_feat, rows, c, vote_xyz,
               vote_feat, norm);
  } else {
    auto kern = vote_tail_fwd_kernel<float>;
    P2R_LAUNCH(kern, grid, 256, 0, st, (const float*)net, seed_xyz, xyz_stride, seed_feat, rows, c, vote_xyz, vote_feat,
               norm);
  }
  P2R_RETURN_LAUNCH("p2r_vote_tail");
}

extern "C" int p2r_vote_tail_grad(const float* g_xyz, const float* g_feat, const float* vote_feat, const float* norm,
                                  long long rows, int c, void* d_net, int net_bf16, float* d_seed_feat, void* stream) {
  P2R_CHECK_ARG(rows >= 0 && c > 0, "p2r_vote_tail_grad");
  if (rows == 0) return 0;
  const int grid = vote_grid(rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (net_bf16) {
    auto kern = vote_tail_bwd_kernel<__nv_bfloat16>;
    P2R_LAUNCH(kern, grid, 256, 0, st, g_xyz, g_feat, vote_feat, norm, rows, c, (__nv_bfloat16*)d_net, d_seed_feat);
  } else {
    auto kern = vote_tail_bwd_kernel<float>;
    P2R_LAUNCH(kern, grid, 256, 0, st, g_xyz, g_feat, vote_feat, norm, rows, c, (float*)d_net, d_seed_feat);
  }
  P2R_RETURN_LAUNCH("p2r_vote_tail_grad");
}
