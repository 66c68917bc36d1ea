// pointnet2_ops.cu -- sm_100a kernels for the nine PointNet++ operators of the reference's
// `pointnet2_ops._ext` ABI (external/pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19),
// re-designed for B200; results are index-exact with the reference kernels
// (sampling_gpu.cu, ball_query_gpu.cu, group_points_gpu.cu, interpolate_gpu.cu).
//
// Design notes (vs the reference's one-CTA-per-batch scalar kernels):
//  * FPS: the cloud lives in REGISTERS for the whole kernel (reference re-reads it from global in
//    each of the m-1 rounds); per round one warp-shuffle arg-max + ONE __syncthreads (double-
//    buffered smem slots); the reference's tie order (its halving tree makes equal distances meet at
//    the LOWEST differing bit of their thread ids, so the smallest BIT-REVERSED (k mod bs) wins, then
//    the smallest k) is encoded in a 64-bit sort key so the result is bit-identical.
//  * ball_query: one WARP per centre scanning 32 candidates per step in index order with
//    ballot/popc compaction, the cloud staged in shared memory by 1-D bulk TMA (UBLKCP);
//    the reference uses one THREAD per centre with uncoalesced 12-byte global loads.
//  * gather/group: grid spread over (points, channel chunks, batch), idx loaded once per thread and
//    reused across channels, 128-bit stores.
//  * three_nn: known cloud staged in smem tiles (bulk TMA), one thread per unknown point.
#include "p2r_common.cuh"

// ------------------------------------------------------------------------------------------------
// reference launch geometry (include/cuda_utils.h:15-19): needed only for the FPS tie order.
static inline int ref_opt_n_threads(int work_size) {
  int pow_2 = 0;
  while ((1 << (pow_2 + 1)) <= work_size) ++pow_2;
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

// ================================================================================================
// furthest point sampling  (replaces sampling_gpu.cu:70-173 + sampling.cpp:66-87)
// ================================================================================================
// key = (float bits of d2) << 32 | ~rank, rank = bitrev(k % bs_ref) * cpt + k / bs_ref  (smaller rank wins)
// key 0 = "no candidate" (the reference's best=-1, besti=0  -> index 0).
__device__ __forceinline__ unsigned fps_rank(int k, int bs_ref, int log2bs, int cpt) {
  const unsigned t = (unsigned)k & (unsigned)(bs_ref - 1);
  const unsigned rev = log2bs == 0 ? 0u : (__brev(t) >> (32 - log2bs));
  return rev * (unsigned)cpt + ((unsigned)k >> log2bs);
}
__device__ __forceinline__ int fps_unrank(unsigned rank, int bs_ref, int log2bs, int cpt) {
  const unsigned rev = rank / (unsigned)cpt, q = rank % (unsigned)cpt;
  const unsigned t = log2bs == 0 ? 0u : (__brev(rev) >> (32 - log2bs));
  return (int)(q * (unsigned)bs_ref + t);
}

// `staged`: the cloud is also kept in (dynamic) shared memory so that the coordinates of each round's pick are read
// from there -- the m - 1 rounds are a dependent chain, and the global load of the pick (L2 latency, ~0.3 us) was the
// longest link of it (round 2: 63 -> see DESIGN.md section 5).  Same values, same order, same result.
// Warp-wide maximum of a 64-bit key with two redux.sync (REDUX.MAX.U32) instead of five shuffle stages of two SHFLs each:
// first the high words (the distance bits: monotone as unsigned for d2 >= 0), then the low words (~rank) among the lanes
// that hold the winning high word.  The rounds of FPS are one dependent chain, so every cycle here is paid m - 1 times.
__device__ __forceinline__ unsigned long long fps_warp_max(unsigned long long key) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  return ((unsigned long long)mhi << 32) | mlo;
}

template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(int n, int m, int bs_ref, int log2bs, int cpt, const float* __restrict__ xyz, int* __restrict__ idxs,
           int staged) {
  constexpr int NW = THREADS / 32;
  __shared__ unsigned long long s_key[2][NW];
  P2R_DYN_SMEM(float, s_pts);    // [n * 3] when staged
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = xyz + (size_t)blockIdx.x * n * 3;
  int* out = idxs + (size_t)blockIdx.x * m;

  float px[PPT], py[PPT], pz[PPT], temp[PPT];
  unsigned nrank[PPT];  // ~rank, 0 => point can never be selected (out of range or |p|^2 <= 1e-3)
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + i * THREADS;
    px[i] = py[i] = pz[i] = 0.f;
    temp[i] = 1e10f;
    nrank[i] = 0u;
    if (k < n) {
      px[i] = __ldg(pts + 3 * k + 0);
      py[i] = __ldg(pts + 3 * k + 1);
      pz[i] = __ldg(pts + 3 * k + 2);
      if (staged) {
        s_pts[3 * k + 0] = px[i];
        s_pts[3 * k + 1] = py[i];
        s_pts[3 * k + 2] = pz[i];
      }
      const float mag = p2r_sqnorm3(px[i], py[i], pz[i]);
      if (!((double)mag <= 1e-3)) {  // sampling_gpu.cu:100-101 (float promoted against a double literal)
        nrank[i] = ~fps_rank(k, bs_ref, log2bs, cpt);
      }
    }
  }
  if (tid == 0) out[0] = 0;
  float x1 = __ldg(pts + 0), y1 = __ldg(pts + 1), z1 = __ldg(pts + 2);
  const float* cpts = staged ? s_pts : pts;
  if (staged) __syncthreads();

  for (int j = 1; j < m; ++j) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float d = p2r_sqdist3(px[i], py[i], pz[i], x1, y1, z1);
      const float d2 = fminf(d, temp[i]);
      if (nrank[i] != 0u) {
        temp[i] = d2;
        const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | nrank[i];
        best = key > best ? key : best;
      }
    }
    best = fps_warp_max(best);
    unsigned long long v;
    if (NW == 1) {
      v = best;
    } else {
      if (lane == 0) s_key[j & 1][warp] = best;
      __syncthreads();
      if (NW <= 8) {               // every thread folds the few per-warp maxima itself: no second shuffle stage
        v = s_key[j & 1][0];
#pragma unroll
        for (int w8 = 1; w8 < NW; ++w8) {
          const unsigned long long o = s_key[j & 1][w8];
          v = o > v ? o : v;
        }
      } else {
        v = fps_warp_max(lane < NW ? s_key[j & 1][lane] : 0ull);
      }
    }
    int old = 0;
    if (v != 0ull) {
      const unsigned rank = ~(unsigned)(v & 0xffffffffull);
      // (cpt == 1 whenever n is at most the reference's block size: rank = bit-reversed index, no division)
      old = cpt == 1 ? (int)(log2bs == 0 ? 0u : (__brev(rank) >> (32 - log2bs))) : fps_unrank(rank, bs_ref, log2bs, cpt);
    }
    if (tid == 0) out[j] = old;
    x1 = cpts[3 * old + 0];
    y1 = cpts[3 * old + 1];
    z1 = cpts[3 * old + 2];
  }
}

// Generic fall-back for clouds that do not fit the register-resident variants (n > 32768):
// same arithmetic and tie order, distances kept in global scratch.
__global__ void __launch_bounds__(1024, 1)
fps_kernel_large(int n, int m, int bs_ref, int log2bs, int cpt, const float* __restrict__ xyz, float* __restrict__ temp_g,
                 int* __restrict__ idxs) {
  __shared__ unsigned long long s_key[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = xyz + (size_t)blockIdx.x * n * 3;
  float* temp = temp_g + (size_t)blockIdx.x * n;
  int* out = idxs + (size_t)blockIdx.x * m;
  for (int k = tid; k < n; k += 1024) temp[k] = 1e10f;
  if (tid == 0) out[0] = 0;
  int old = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = __ldg(pts + 3 * old), y1 = __ldg(pts + 3 * old + 1), z1 = __ldg(pts + 3 * old + 2);
    unsigned long long best = 0ull;
    for (int k = tid; k < n; k += 1024) {
      const float x2 = __ldg(pts + 3 * k), y2 = __ldg(pts + 3 * k + 1), z2 = __ldg(pts + 3 * k + 2);
      const float mag = p2r_sqnorm3(x2, y2, z2);
      if ((double)mag <= 1e-3) continue;
      const float d2 = fminf(p2r_sqdist3(x2, y2, z2, x1, y1, z1), temp[k]);
      temp[k] = d2;
      const unsigned long long key =
          ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(~fps_rank(k, bs_ref, log2bs, cpt));
      best = key > best ? key : best;
    }
    for (int off = 16; off >= 1; off >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
      best = o > best ? o : best;
    }
    if (lane == 0) s_key[j & 1][warp] = best;
    __syncthreads();
    unsigned long long v = s_key[j & 1][lane];
    for (int off = 16; off >= 1; off >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, off);
      v = o > v ? o : v;
    }
    old = 0;
    if (v != 0ull) {
      old = fps_unrank(~(unsigned)(v & 0xffffffffull), bs_ref, log2bs, cpt);
    }
    if (tid == 0) out[j] = old;
  }
}

template <int PPT, int THREADS>
static void launch_fps(int b, int n, int m, int bs_ref, int cpt, const float* xyz, int* idxs, cudaStream_t st) {
  int log2bs = 0;
  while ((1 << log2bs) < bs_ref) ++log2bs;
  auto kern = fps_kernel<PPT, THREADS>;
  const int staged = (size_t)n * 12 <= 48 * 1024 ? 1 : 0;      // clouds of up to 4096 points ride in shared memory
  P2R_LAUNCH(kern, b, THREADS, staged ? (size_t)n * 12 : 0, st, n, m, bs_ref, log2bs, cpt, xyz, idxs, staged);
}

extern "C" int p2r_furthest_point_sampling(const float* xyz, int b, int n, int m, int* idxs, float* scratch,
                                           void* stream) {
  P2R_CHECK_ARG(b >= 0 && n > 0 && m >= 0, "p2r_furthest_point_sampling");
  if (b == 0 || m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int bs_ref = ref_opt_n_threads(n);
  const int cpt = (n + bs_ref - 1) / bs_ref;
  if (n <= 128) launch_fps<1, 128>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 256) launch_fps<2, 128>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 512) launch_fps<4, 128>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 1024) launch_fps<4, 256>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 2048) launch_fps<4, 512>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 4096) launch_fps<4, 1024>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 8192) launch_fps<8, 1024>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 16384) launch_fps<16, 1024>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else if (n <= 32768) launch_fps<32, 1024>(b, n, m, bs_ref, cpt, xyz, idxs, st);
  else {
    P2R_CHECK_ARG(scratch != nullptr, "p2r_furthest_point_sampling (n > 32768 needs b*n floats of scratch)");
    int log2bs = 0;
    while ((1 << log2bs) < bs_ref) ++log2bs;
    P2R_LAUNCH(fps_kernel_large, b, 1024, 0, st, n, m, bs_ref, log2bs, cpt, xyz, scratch, idxs);
  }
  P2R_RETURN_LAUNCH("p2r_furthest_point_sampling");
}

// ================================================================================================
// gather_points (+grad)   (replaces sampling_gpu.cu:8-47)      points (B,C,N), idx (B,M) -> (B,C,M)
// ================================================================================================
template <bool GRAD>
__global__ void __launch_bounds__(256)
gather_points_kernel(int c, int n, int m, int c_per_cta, const float* __restrict__ src, const int* __restrict__ idx,
                     float* __restrict__ dst) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const int a = __ldg(idx + (size_t)b * m + j);
  const int c0 = blockIdx.y * c_per_cta;
  const int c1 = min(c, c0 + c_per_cta);
  for (int l = c0; l < c1; ++l) {
    if (!GRAD) dst[((size_t)b * c + l) * m + j] = __ldg(src + ((size_t)b * c + l) * n + a);
    else atomicAdd(dst + ((size_t)b * c + l) * n + a, __ldg(src + ((size_t)b * c + l) * m + j));
  }
}

extern "C" int p2r_gather_points(const float* points, const int* idx, int b, int c, int n, int m, float* out,
                                 void* stream) {
  P2R_CHECK_ARG(b >= 0 && c >= 0 && n > 0 && m >= 0, "p2r_gather_points");
  if (b == 0 || c == 0 || m == 0) return 0;
  const int cpc = c >= 64 ? 16 : (c >= 8 ? 4 : 1);
  dim3 grid(p2r_ceil_div(m, 256), p2r_ceil_div(c, cpc), b);
  auto kern = gather_points_kernel<false>;
  P2R_LAUNCH(kern, grid, 256, 0, (cudaStream_t)stream, c, n, m, cpc, points, idx, out);
  P2R_RETURN_LAUNCH("p2r_gather_points");
}

// grad_points (B,C,N) must be zero-filled by the caller (the reference allocates torch::zeros,
// sampling.cpp:49-51); duplicates in idx accumulate.
extern "C" int p2r_gather_points_grad(const float* grad_out, const int* idx, int b, int c, int n, int m,
                                      float* grad_points, void* stream) {
  P2R_CHECK_ARG(b >= 0 && c >= 0 && n > 0 && m >= 0, "p2r_gather_points_grad");
  if (b == 0 || c == 0 || m == 0) return 0;
  const int cpc = c >= 64 ? 16 : (c >= 8 ? 4 : 1);
  dim3 grid(p2r_ceil_div(m, 256), p2r_ceil_div(c, cpc), b);
  auto kern = gather_points_kernel<true>;
  P2R_LAUNCH(kern, grid, 256, 0, (cudaStream_t)stream, c, n, m, cpc, grad_out, idx, grad_points);
  P2R_RETURN_LAUNCH("p2r_gather_points_grad");
}

// ================================================================================================
// ball_query  (replaces ball_query_gpu.cu:9-44 + ball_query.cpp:19-21)
// new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample): first nsample hits in index order with
// d2 < r*r (strict), unfilled slots = first hit, no hit = 0.   One warp per centre.
// ================================================================================================
#define BQ_WARPS 8
#define BQ_TILE 4096  // points per smem tile (48 KB)
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int n, int m, float radius, int nsample, const float* __restrict__ new_xyz,
                  const float* __restrict__ xyz, int* __restrict__ idx) {
  P2R_DYN_SMEM(float, s_pts);
  __shared__ __align__(8) uint64_t s_bar;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * BQ_WARPS + warp;
  const float* pts = xyz + (size_t)b * n * 3;
  const float radius2 = __fmul_rn(radius, radius);
  if (threadIdx.x == 0) {
    p2r_mbar_init(&s_bar, 1);
    p2r_fence_mbar_init();
  }
  __syncthreads();

  float cx = 0.f, cy = 0.f, cz = 0.f;
  const bool active_warp = j < m;
  if (active_warp) {
    const float* q = new_xyz + ((size_t)b * m + j) * 3;
    cx = __ldg(q + 0); cy = __ldg(q + 1); cz = __ldg(q + 2);
  }
  int* o = idx + ((size_t)b * m + (active_warp ? j : 0)) * nsample;
  int cnt = active_warp ? 0 : nsample;  // inactive warps are "done"
  int first = 0;
  uint32_t parity = 0;
  for (int base = 0; base < n; base += BQ_TILE) {
    const int tile = min(BQ_TILE, n - base);
    p2r_stage_floats(s_pts, pts + (size_t)base * 3, tile * 3, &s_bar, parity);
    parity ^= 1u;
    for (int k0 = 0; k0 < tile && cnt < nsample; k0 += 32) {
      const int k = k0 + lane;
      bool hit = false;
      if (k < tile) {
        const float d2 = p2r_sqdist3(cx, cy, cz, s_pts[3 * k], s_pts[3 * k + 1], s_pts[3 * k + 2]);
        hit = d2 < radius2;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        if (cnt == 0) first = base + k0 + (__ffs(bal) - 1);
        const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        if (hit && pos < nsample) o[pos] = base + k;
        cnt += __popc(bal);
      }
    }
    // every warp of the CTA done? (uniform decision; also orders smem reuse for the next tile)
    if (__syncthreads_and(cnt >= nsample)) break;
  }
  if (active_warp) {
    const int filled = min(cnt, nsample);
    const int fill = cnt > 0 ? first : 0;
    for (int l = filled + lane; l < nsample; l += 32) o[l] = fill;
  }
}

extern "C" int p2r_ball_query(const float* new_xyz, const float* xyz, int b, int n, int m, float radius,
                              int nsample, int* idx, void* stream) {
  P2R_CHECK_ARG(b >= 0 && n > 0 && m >= 0 && nsample > 0, "p2r_ball_query");
  if (b == 0 || m == 0) return 0;
  const int tile = n < BQ_TILE ? n : BQ_TILE;
  const size_t smem = (size_t)tile * 3 * sizeof(float) + 16;
  dim3 grid(p2r_ceil_div(m, BQ_WARPS), b);
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  P2R_LAUNCH(ball_query_kernel, grid, BQ_WARPS * 32, smem, (cudaStream_t)stream, n, m, radius, nsample, new_xyz, xyz, idx);
  P2R_RETURN_LAUNCH("p2r_ball_query");
}

// ================================================================================================
// group_points (+grad)  (replaces group_points_gpu.cu:8-64)
// points (B,C,N), idx (B,P,S) -> out (B,C,P,S).  Thread = 4 consecutive (p,s) slots x channel chunk.
// ================================================================================================
__global__ void __launch_bounds__(256)
group_points_kernel(int c, int n, int ps, int c_per_cta, const float* __restrict__ points,
                    const int* __restrict__ idx, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (q >= ps) return;
  const int* ip = idx + (size_t)b * ps + q;
  int ii[4];
  const int valid = min(4, ps - q);
#pragma unroll
  for (int t = 0; t < 4; ++t) ii[t] = t < valid ? __ldg(ip + t) : 0;
  const int c0 = blockIdx.y * c_per_cta, c1 = min(c, c0 + c_per_cta);
  const bool vec = (valid == 4) && ((ps & 3) == 0);
  for (int l = c0; l < c1; ++l) {
    const float* row = points + ((size_t)b * c + l) * n;
    float* dst = out + ((size_t)b * c + l) * ps + q;
    if (vec) {
      float4 v = make_float4(__ldg(row + ii[0]), __ldg(row + ii[1]), __ldg(row + ii[2]), __ldg(row + ii[3]));
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      for (int t = 0; t < valid; ++t) dst[t] = __ldg(row + ii[t]);
    }
  }
}

__global__ void __launch_bounds__(256)
group_points_grad_kernel(int c, int n, int ps, int c_per_cta, const float* __restrict__ grad_out,
                         const int* __restrict__ idx, float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= ps) return;
  const int ii = __ldg(idx + (size_t)b * ps + q);
  const int c0 = blockIdx.y * c_per_cta, c1 = min(c, c0 + c_per_cta);
  for (int l = c0; l < c1; ++l)
    atomicAdd(grad_points + ((size_t)b * c + l) * n + ii, __ldg(grad_out + ((size_t)b * c + l) * ps + q));
}

extern "C" int p2r_group_points(const float* points, const int* idx, int b, int c, int n, int npoints, int nsample,
                                float* out, void* stream) {
  P2R_CHECK_ARG(b >= 0 && c >= 0 && n > 0 && npoints >= 0 && nsample >= 0, "p2r_group_points");
  const int ps = npoints * nsample;
  if (b == 0 || c == 0 || ps == 0) return 0;
  const int cpc = c >= 64 ? 8 : (c >= 8 ? 4 : 1);
  dim3 grid(p2r_ceil_div(p2r_ceil_div(ps, 4), 256), p2r_ceil_div(c, cpc), b);
  P2R_LAUNCH(group_points_kernel, grid, 256, 0, (cudaStream_t)stream, c, n, ps, cpc, points, idx, out);
  P2R_RETURN_LAUNCH("p2r_group_points");
}

// grad_points (B,C,N) must be zero-filled by the caller (group_points.cpp:48-50).
extern "C" int p2r_group_points_grad(const float* grad_out, const int* idx, int b, int c, int n, int npoints,
                                     int nsample, float* grad_points, void* stream) {
  P2R_CHECK_ARG(b >= 0 && c >= 0 && n > 0 && npoints >= 0 && nsample >= 0, "p2r_group_points_grad");
  const int ps = npoints * nsample;
  if (b == 0 || c == 0 || ps == 0) return 0;
  const int cpc = c >= 64 ? 8 : (c >= 8 ? 4 : 1);
  dim3 grid(p2r_ceil_div(ps, 256), p2r_ceil_div(c, cpc), b);
  P2R_LAUNCH(group_points_grad_kernel, grid, 256, 0, (cudaStream_t)stream, c, n, ps, cpc, grad_out, idx, grad_points);
  P2R_RETURN_LAUNCH("p2r_group_points_grad");
}

// ================================================================================================
// three_nn  (replaces interpolate_gpu.cu:9-59)  unknown (B,n,3), known (B,m,3) -> dist2, idx (B,n,3)
// The reference compares a float candidate against double bests initialised to 1e40 with strict <;
// float bests initialised to +inf give the same decisions and the same outputs ((float)1e40 = inf).
// ================================================================================================
#define TNN_TILE 2048
__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ __align__(16) float s_known[TNN_TILE * 3];
  __shared__ __align__(8) uint64_t s_bar;
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (threadIdx.x == 0) {
    p2r_mbar_init(&s_bar, 1);
    p2r_fence_mbar_init();
  }
  __syncthreads();
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (j < n) {
    const float* u = unknown + ((size_t)b * n + j) * 3;
    ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
  }
  const float inf = __int_as_float(0x7f800000);
  float b1 = inf, b2 = inf, b3 = inf;
  int i1 = 0, i2 = 0, i3 = 0;
  uint32_t parity = 0;
  for (int base = 0; base < m; base += TNN_TILE) {
    const int tile = min(TNN_TILE, m - base);
    if (base > 0) __syncthreads();
    p2r_stage_floats(s_known, known + ((size_t)b * m + base) * 3, tile * 3, &s_bar, parity);
    parity ^= 1u;
    for (int k = 0; k < tile; ++k) {
      const float d = p2r_sqdist3(ux, uy, uz, s_known[3 * k], s_known[3 * k + 1], s_known[3 * k + 2]);
      const int kk = base + k;
      if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk; }
      else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = kk; }
      else if (d < b3) { b3 = d; i3 = kk; }
    }
  }
  if (j < n) {
    const size_t o = ((size_t)b * n + j) * 3;
    dist2[o] = b1; dist2[o + 1] = b2; dist2[o + 2] = b3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
  }
}

extern "C" int p2r_three_nn(const float* unknown, const float* known, int b, int n, int m, float* dist2, int* idx,
                            void* stream) {
  P2R_CHECK_ARG(b >= 0 && n >= 0 && m >= 0, "p2r_three_nn");
  if (b == 0 || n == 0) return 0;
  dim3 grid(p2r_ceil_div(n, 256), b);
  P2R_LAUNCH(three_nn_kernel, grid, 256, 0, (cudaStream_t)stream, n, m, unknown, known, dist2, idx);
  P2R_RETURN_LAUNCH("p2r_three_nn");
}

// ================================================================================================
// three_interpolate (+grad)  (replaces interpolate_gpu.cu:72-143)
// points (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n)
// ================================================================================================
template <bool GRAD>
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, int c_per_cta, const float* __restrict__ src,
                         const int* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ dst) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t o = ((size_t)b * n + j) * 3;
  const float w1 = __ldg(weight + o), w2 = __ldg(weight + o + 1), w3 = __ldg(weight + o + 2);
  const int i1 = __ldg(idx + o), i2 = __ldg(idx + o + 1), i3 = __ldg(idx + o + 2);
  const int c0 = blockIdx.y * c_per_cta, c1 = min(c, c0 + c_per_cta);
  for (int l = c0; l < c1; ++l) {
    if (!GRAD) {
      const float* p = src + ((size_t)b * c + l) * m;
      float t = __fmul_rn(__ldg(p + i2), w2);  // reference SASS order: p2*w2, fma(p1,w1,.), fma(p3,w3,.)
      t = __fmaf_rn(__ldg(p + i1), w1, t);
      t = __fmaf_rn(__ldg(p + i3), w3, t);
      dst[((size_t)b * c + l) * n + j] = t;
    } else {
      const float g = __ldg(src + ((size_t)b * c + l) * n + j);
      float* gp = dst + ((size_t)b * c + l) * m;
      atomicAdd(gp + i1, __fmul_rn(g, w1));
      atomicAdd(gp + i2, __fmul_rn(g, w2));
      atomicAdd(gp + i3, __fmul_rn(g, w3));
    }
  }
}

extern "C" int p2r_three_interpolate(const float* points, const int* idx, const float* weight, int b, int c, int m,
                                     int n, float* out, void* stream) {
  P2R_CHECK_ARG(b >= 0 && c >= 0 && m > 0 && n >= 0, "p2r_three_interpolate");
  if (b == 0 || c == 0 || n == 0) return 0;
  const int cpc = c >= 64 ? 8 : (c >= 8 ? 4 : 1);
  dim3 grid(p2r_ceil_div(n, 256), p2r_ceil_div(c, cpc), b);
  auto kern = three_interpolate_kernel<false>;
  P2R_LAUNCH(kern, grid, 256, 0, (cudaStream_t)stream, c, m, n, cpc, points, idx, weight, out);
  P2R_RETURN_LAUNCH("p2r_three_interpolate");
}

// grad_points (B,C,m) must be zero-filled by the caller (interpolate.cpp:85-87).
extern "C" int p2r_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight, int b, int c,
                                          int n, int m, float* grad_points, void* stream) {
  P2R_CHECK_ARG(b >= 0 && c >= 0 && m > 0 && n >= 0, "p2r_three_interpolate_grad");
  if (b == 0 || c == 0 || n == 0) return 0;
  const int cpc = c >= 64 ? 8 : (c >= 8 ? 4 : 1);
  dim3 grid(p2r_ceil_div(n, 256), p2r_ceil_div(c, cpc), b);
  auto kern = three_interpolate_kernel<true>;
  P2R_LAUNCH(kern, grid, 256, 0, (cudaStream_t)stream, c, m, n, cpc, grad_out, idx, weight, grad_points);
  P2R_RETURN_LAUNCH("p2r_three_interpolate_grad");
}
