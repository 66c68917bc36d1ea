// geometry_ops.cu -- sm_100a kernels for the graph / loss / eval geometry of the P2RNet hot path:
//   knn_graph            net_utils/vn_dgcnn_util.py:4-10   (top-k of -|xi-xj|^2, self included)
//   graph_offset         net_utils/vn_dgcnn_util.py:70-95  (x[idx] - x)
//   nn_distance (+grad)  net_utils/nn_distance.py:34-61    (all-pairs cost, row/col min + argmin)
//   decode_boxes         net_utils/ap_helper.py:152-196    (exp/atan2 decode, 8 corners, far-box test, AABB)
//   nms3d                net_utils/nms.py:41-119           (greedy AABB NMS, selection-exact)
//   box3d_iou            net_utils/box_util.py:90-118      (oriented-box IoU via polygon clipping)
// All citations are into /root/reference.
#include "p2r_common.cuh"
#include <math.h>

// ================================================================================================
// knn_graph: x (B,C,N) f32 -> idx (B,N,k) i64, sorted by descending -(squared distance);
// arithmetic order fixed as: xx_i = sum_c x_ci*x_ci (sequential adds of rounded squares),
// s_ij = x_0i*x_0j then fma over c, inner = -2*s, pd = (-xx_i - inner) - xx_j; ties -> lower j.
// One thread per query row, candidates staged in shared memory tiles.
// ================================================================================================
#define KNN_MAXK 64
#define KNN_TILE 256
__global__ void __launch_bounds__(128)
knn_kernel(int c, int n, int k, const float* __restrict__ x, long long* __restrict__ idx) {
  P2R_DYN_SMEM(float, s_x);  // [c][KNN_TILE] + xx[KNN_TILE]
  float* s_xx = s_x + (size_t)c * KNN_TILE;
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* xb = x + (size_t)b * c * n;
  float best_v[KNN_MAXK];
  int best_i[KNN_MAXK];
  int filled = 0;
  float xxi = 0.f;
  if (i < n) {
    for (int cc = 0; cc < c; ++cc) {
      const float v = __ldg(xb + (size_t)cc * n + i);
      const float sq = __fmul_rn(v, v);
      xxi = cc == 0 ? sq : __fadd_rn(xxi, sq);
    }
  }
  for (int base = 0; base < n; base += KNN_TILE) {
    const int tile = min(KNN_TILE, n - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tile; t += blockDim.x) {
      float xx = 0.f;
      for (int cc = 0; cc < c; ++cc) {
        const float v = __ldg(xb + (size_t)cc * n + base + t);
        s_x[cc * KNN_TILE + t] = v;
        const float sq = __fmul_rn(v, v);
        xx = cc == 0 ? sq : __fadd_rn(xx, sq);
      }
      s_xx[t] = xx;
    }
    __syncthreads();
    if (i < n) {
      for (int t = 0; t < tile; ++t) {
        float s = 0.f;
        for (int cc = 0; cc < c; ++cc) {
          const float a = __ldg(xb + (size_t)cc * n + i);
          s = cc == 0 ? __fmul_rn(a, s_x[t]) : __fmaf_rn(a, s_x[cc * KNN_TILE + t], s);
        }
        const float inner = __fmul_rn(-2.f, s);
        const float pd = __fsub_rn(__fsub_rn(-xxi, inner), s_xx[t]);
        // insert into the descending list; strict > keeps the earlier index ahead on ties
        if (filled < k || pd > best_v[filled - 1]) {
          int pos = filled < k ? filled : k - 1;
          while (pos > 0 && pd > best_v[pos - 1]) {
            best_v[pos] = best_v[pos - 1];
            best_i[pos] = best_i[pos - 1];
            --pos;
          }
          best_v[pos] = pd;
          best_i[pos] = base + t;
          if (filled < k) ++filled;
        }
      }
    }
  }
  if (i < n) {
    long long* o = idx + ((size_t)b * n + i) * k;
    for (int t = 0; t < k; ++t) o[t] = t < filled ? best_i[t] : 0;
  }
}

extern "C" int p2r_knn_graph(const float* x, int b, int c, int n, int k, long long* idx, void* stream) {
  P2R_CHECK_ARG(b >= 0 && c > 0 && n > 0 && k > 0 && k <= KNN_MAXK && k <= n, "p2r_knn_graph");
  if (b == 0) return 0;
  const size_t smem = ((size_t)c + 1) * KNN_TILE * sizeof(float);
  P2R_CHECK_ARG(smem <= 200 * 1024, "p2r_knn_graph (feature dim too large for the smem tile)");
  if (smem > 48 * 1024) cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(p2r_ceil_div(n, 128), b);
  P2R_LAUNCH(knn_kernel, grid, 128, smem, (cudaStream_t)stream, c, n, k, x, idx);
  P2R_RETURN_LAUNCH("p2r_knn_graph");
}

// graph_offset: x (B,D3,N) with D3 = 3*d, idx (B,N,k) i64 -> out (B,N,k,d,3) = x[:, :, idx] - x
__global__ void __launch_bounds__(256)
graph_offset_kernel(int d3, int n, int k, const float* __restrict__ x, const long long* __restrict__ idx,
                    float* __restrict__ out, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int ch = (int)(e % d3);
  const long long r = e / d3;  // (b*n + i)*k + t
  const int i = (int)((r / k) % n);
  const int b = (int)(r / ((long long)k * n));
  const int j = (int)__ldg(idx + r);
  const float* xb = x + ((size_t)b * d3 + ch) * n;
  out[e] = __fsub_rn(__ldg(xb + j), __ldg(xb + i));
}

extern "C" int p2r_graph_offset(const float* x, const long long* idx, int b, int d3, int n, int k, float* out,
                                void* stream) {
  P2R_CHECK_ARG(b >= 0 && d3 > 0 && n > 0 && k > 0, "p2r_graph_offset");
  const long long total = (long long)b * n * k * d3;
  if (total == 0) return 0;
  P2R_LAUNCH(graph_offset_kernel, p2r_ceil_div(total, 256), 256, 0, (cudaStream_t)stream, d3, n, k, x, idx, out, total);
  P2R_RETURN_LAUNCH("p2r_graph_offset");
}


// ================================================================================================
// uniform_seed_inds: arc-length-uniform frame sampling (models/p2rnet/modules/stgcn.py:96-101).
// hip (B,T,3) with `stride` floats between frames -> seed_inds (B,S) i64 =
//   argmin_t | cum[t] - s * cum[T-1]/(S-1) |, cum = cumulative hip path length.
// Arithmetic mirrors the reference as executed by torch on CPU (what the golden vectors pin):
// step = sqrtf(fma(dz,dz,fma(dy,dy,dx*dx))), cumulative sum accumulated in double and rounded to float
// per element, fp32 division / multiply / subtract, first minimal index wins.
// One CTA per sequence; T <= 16384.
// ================================================================================================
__global__ void __launch_bounds__(256)
uniform_seed_kernel(int T, int S, int stride, const float* __restrict__ hip, long long* __restrict__ seed) {
  P2R_DYN_SMEM(float, s_cum);  // [T]
  const int b = blockIdx.x;
  const float* h = hip + (size_t)b * T * stride;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float d = 0.f;
    if (t > 0) {
      const float dx = __fsub_rn(__ldg(h + (size_t)t * stride), __ldg(h + (size_t)(t - 1) * stride));
      const float dy = __fsub_rn(__ldg(h + (size_t)t * stride + 1), __ldg(h + (size_t)(t - 1) * stride + 1));
      const float dz = __fsub_rn(__ldg(h + (size_t)t * stride + 2), __ldg(h + (size_t)(t - 1) * stride + 2));
      d = __fsqrt_rn(p2r_sqnorm3_xyz(dx, dy, dz));
    }
    s_cum[t] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int t = 0; t < T; ++t) {
      acc += (double)s_cum[t];
      s_cum[t] = (float)acc;
    }
  }
  __syncthreads();
  const float step = __fdiv_rn(s_cum[T - 1], (float)(S - 1));
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float target = __fmul_rn(step, (float)s);
    float best = fabsf(__fsub_rn(s_cum[0], target));
    int bi = 0;
    for (int t = 1; t < T; ++t) {
      const float v = fabsf(__fsub_rn(s_cum[t], target));
      if (v < best) { best = v; bi = t; }
    }
    seed[(size_t)b * S + s] = bi;
  }
}

extern "C" int p2r_uniform_seed_inds(const float* hip, int stride, int b, int t, int s, long long* seed_inds,
                                     void* stream) {
  P2R_CHECK_ARG(b >= 0 && t > 0 && t <= 16384 && s > 1 && stride >= 3, "p2r_uniform_seed_inds");
  if (b == 0) return 0;
  const size_t smem = (size_t)t * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(uniform_seed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  P2R_LAUNCH(uniform_seed_kernel, b, 256, smem, (cudaStream_t)stream, t, s, stride, hip, seed_inds);
  P2R_RETURN_LAUNCH("p2r_uniform_seed_inds");
}

// ================================================================================================
// nn_distance: pc1 (B,N,C), pc2 (B,M,C) -> dist1,idx1 (B,N), dist2,idx2 (B,M).
// mode 0: squared L2, 1: L1, 2: smooth-L1 (huber, delta).  Per-pair cost = sequential fp32 sum over
// C of separately rounded terms (torch materialises diff**2 before summing: no FMA); first minimal
// index wins (torch.min).  grid.x tiles the "row" cloud; blockIdx.z = 0 -> rows of pc1, 1 -> pc2.
// ================================================================================================
#define NND_TILE 1024
__device__ __forceinline__ float nnd_term(float a, float b, int mode, float delta) {
  const float diff = __fsub_rn(a, b);
  if (mode == 0) return __fmul_rn(diff, diff);
  const float ab = fabsf(diff);
  if (mode == 1) return ab;
  const float q = fminf(ab, delta);
  const float lin = __fsub_rn(ab, q);
  return __fadd_rn(__fmul_rn(0.5f, __fmul_rn(q, q)), __fmul_rn(delta, lin));
}

__global__ void __launch_bounds__(128)
nn_distance_kernel(int n, int m, int c, int mode, float delta, const float* __restrict__ pc1,
                   const float* __restrict__ pc2, float* __restrict__ dist1, long long* __restrict__ idx1,
                   float* __restrict__ dist2, long long* __restrict__ idx2) {
  P2R_DYN_SMEM(float, s_other);  // [tile][c]
  const int b = blockIdx.x;      // the batch runs along grid.x: the vote loss calls this with B * num_seeds "batches"
  const bool second = blockIdx.z == 1;      // (grid.y is limited to 65535, grid.x is not)
  const int nr = second ? m : n, no = second ? n : m;
  const float* rows = (second ? pc2 : pc1) + (size_t)b * nr * c;
  const float* others = (second ? pc1 : pc2) + (size_t)b * no * c;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if ((long long)blockIdx.y * blockDim.x >= nr) return;  // whole CTA out of range (uniform)
  float best = __int_as_float(0x7f800000);
  int besti = 0;
  bool any = false;
  const int tile_cap = NND_TILE;
  for (int base = 0; base < no; base += tile_cap) {
    const int tile = min(tile_cap, no - base);
    __syncthreads();
    for (int t = threadIdx.x; t < tile * c; t += blockDim.x) s_other[t] = __ldg(others + (size_t)base * c + t);
    __syncthreads();
    if (i < nr) {
      for (int t = 0; t < tile; ++t) {
        float s = 0.f;
        for (int cc = 0; cc < c; ++cc) {
          const float term = nnd_term(second ? s_other[t * c + cc] : __ldg(rows + (size_t)i * c + cc),
                                      second ? __ldg(rows + (size_t)i * c + cc) : s_other[t * c + cc], mode, delta);
          s = cc == 0 ? term : __fadd_rn(s, term);
        }
        // torch.min: first minimal index; NaN propagates in torch, ignored here (inputs are finite)
        if (!any || s < best) { best = s; besti = base + t; any = true; }
      }
    }
  }
  if (i < nr) {
    (second ? dist2 : dist1)[(size_t)b * nr + i] = best;
    (second ? idx2 : idx1)[(size_t)b * nr + i] = besti;
  }
}

extern "C" int p2r_nn_distance(const float* pc1, const float* pc2, int b, int n, int m, int c, int mode, float delta,
                               float* dist1, long long* idx1, float* dist2, long long* idx2, void* stream) {
  P2R_CHECK_ARG(b >= 0 && n > 0 && m > 0 && c > 0 && mode >= 0 && mode <= 2, "p2r_nn_distance");
  if (b == 0) return 0;
  const int mx = n > m ? n : m;
  const size_t smem = (size_t)(mx < NND_TILE ? mx : NND_TILE) * c * sizeof(float);
  P2R_CHECK_ARG(smem <= 200 * 1024, "p2r_nn_distance (C too large for the smem tile)");
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(nn_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  P2R_CHECK_ARG(p2r_ceil_div(mx, 128) <= 65535, "p2r_nn_distance (more than 8 M points per cloud)");
  dim3 grid(b, p2r_ceil_div(mx, 128), 2);
  P2R_LAUNCH(nn_distance_kernel, grid, 128, smem, (cudaStream_t)stream, n, m, c, mode, delta, pc1, pc2, dist1, idx1, dist2, idx2);
  P2R_RETURN_LAUNCH("p2r_nn_distance");
}

// backward: d dist1[b,i] / d pc1[b,i,:] and d pc2[b,idx1[b,i],:] (and symmetrically for dist2);
// grad_pc1 / grad_pc2 must be zero-filled by the caller.
__device__ __forceinline__ float nnd_dterm(float a, float b, int mode, float delta) {
  const float diff = a - b;
  if (mode == 0) return 2.f * diff;
  const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
  if (mode == 1) return sg;
  return fabsf(diff) <= delta ? diff : delta * sg;
}

__global__ void __launch_bounds__(256)
nn_distance_grad_kernel(int n, int m, int c, int mode, float delta, const float* __restrict__ pc1,
                        const float* __restrict__ pc2, const long long* __restrict__ idx1,
                        const long long* __restrict__ idx2, const float* __restrict__ g1,
                        const float* __restrict__ g2, float* __restrict__ grad_pc1, float* __restrict__ grad_pc2,
                        int b) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long t1 = (long long)b * n * c, t2 = (long long)b * m * c;
  if (e < t1) {
    const int cc = (int)(e % c);
    const long long r = e / c;  // b*n + i
    const int bb = (int)(r / n);
    const int j = (int)__ldg(idx1 + r);
    const float g = g1 ? __ldg(g1 + r) : 0.f;
    if (g != 0.f) {
      const float d = g * nnd_dterm(__ldg(pc1 + e), __ldg(pc2 + ((size_t)bb * m + j) * c + cc), mode, delta);
      atomicAdd(grad_pc1 + e, d);
      atomicAdd(grad_pc2 + ((size_t)bb * m + j) * c + cc, -d);
    }
  } else if (e < t1 + t2) {
    const long long e2 = e - t1;
    const int cc = (int)(e2 % c);
    const long long r = e2 / c;  // b*m + j
    const int bb = (int)(r / m);
    const int i = (int)__ldg(idx2 + r);
    const float g = g2 ? __ldg(g2 + r) : 0.f;
    if (g != 0.f) {
      const float d = g * nnd_dterm(__ldg(pc1 + ((size_t)bb * n + i) * c + cc), __ldg(pc2 + e2), mode, delta);
      atomicAdd(grad_pc1 + ((size_t)bb * n + i) * c + cc, d);
      atomicAdd(grad_pc2 + e2, -d);
    }
  }
}

extern "C" int p2r_nn_distance_grad(const float* pc1, const float* pc2, const long long* idx1, const long long* idx2,
                                    const float* g1, const float* g2, int b, int n, int m, int c, int mode,
                                    float delta, float* grad_pc1, float* grad_pc2, void* stream) {
  P2R_CHECK_ARG(b >= 0 && n > 0 && m > 0 && c > 0 && mode >= 0 && mode <= 2, "p2r_nn_distance_grad");
  const long long total = (long long)b * (n + m) * c;
  if (total == 0) return 0;
  P2R_LAUNCH(nn_distance_grad_kernel, p2r_ceil_div(total, 256), 256, 0, (cudaStream_t)stream, n, m, c, mode, delta, pc1, pc2, idx1, idx2, g1, g2, grad_pc1, grad_pc2, b);
  P2R_RETURN_LAUNCH("p2r_nn_distance_grad");
}

// ================================================================================================
// decode_boxes: network outputs -> 8 oriented corners (f64), axis-aligned hull (f64), far-box mask.
// Mirrors ap_helper.py:152-196: size = exp(log_size) in fp32, theta = atan2(sin, cos) in f64,
// vectors = diag(size/2 [fp32]) . head2rot(theta) in f64, corners per utils/tools.py:33-51;
// a proposal is "empty" when any size < 0.01 or > 10, or when no hip point lies inside the box
// enlarged by `contact` (analytic |R(p-c)| <= size/2 + contact instead of scipy Delaunay).
// One warp per proposal.
// ================================================================================================
__global__ void __launch_bounds__(128)
decode_boxes_kernel(int k, int t, const float* __restrict__ center, const float* __restrict__ log_size,
                    const double* __restrict__ heading, const float* __restrict__ hip, int hip_stride,
                    double contact, double* __restrict__ corners, double* __restrict__ aabb,
                    unsigned char* __restrict__ nonempty, int total) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= total) return;
  const int b = w / k;
  const float sx = expf(__ldg(log_size + (size_t)w * 3 + 0));
  const float sy = expf(__ldg(log_size + (size_t)w * 3 + 1));
  const float sz = expf(__ldg(log_size + (size_t)w * 3 + 2));
  const double th = atan2(__ldg(heading + (size_t)w * 2 + 0), __ldg(heading + (size_t)w * 2 + 1));
  const double cs = cos(th), sn = sin(th);
  const double cx = (double)__ldg(center + (size_t)w * 3 + 0), cy = (double)__ldg(center + (size_t)w * 3 + 1),
               cz = (double)__ldg(center + (size_t)w * 3 + 2);
  // half extents: numpy computes float32 size / 2. in float32, then promotes
  const double hx = (double)(sx * 0.5f), hy = (double)(sy * 0.5f), hz = (double)(sz * 0.5f);
  // rows of diag(h) . R, R = [[c,0,-s],[0,1,0],[s,0,c]]
  const double v0[3] = {__dmul_rn(hx, cs), 0.0, __dmul_rn(hx, -sn)};
  const double v1[3] = {0.0, hy, 0.0};
  const double v2[3] = {__dmul_rn(hz, sn), 0.0, __dmul_rn(hz, cs)};
  if (lane < 8) {
    const double s0 = (lane == 1 || lane == 2 || lane == 5 || lane == 6) ? 1.0 : -1.0;
    const double s1 = (lane == 2 || lane == 3 || lane == 6 || lane == 7) ? 1.0 : -1.0;
    const double s2 = lane >= 4 ? 1.0 : -1.0;
    const double c3[3] = {cx, cy, cz};
    for (int a = 0; a < 3; ++a) {
      double p = __dadd_rn(c3[a], s0 * v0[a]);
      p = __dadd_rn(p, s1 * v1[a]);
      p = __dadd_rn(p, s2 * v2[a]);
      corners[((size_t)w * 8 + lane) * 3 + a] = p;
    }
  }
  // axis-aligned hull of the 8 corners (min / max over lanes 0..7)
  for (int a = 0; a < 3; ++a) {
    double p = 0.0;
    {
      const int l = lane & 7;
      const double s0 = (l == 1 || l == 2 || l == 5 || l == 6) ? 1.0 : -1.0;
      const double s1 = (l == 2 || l == 3 || l == 6 || l == 7) ? 1.0 : -1.0;
      const double s2 = l >= 4 ? 1.0 : -1.0;
      const double c3 = a == 0 ? cx : (a == 1 ? cy : cz);
      p = __dadd_rn(c3, s0 * v0[a]);
      p = __dadd_rn(p, s1 * v1[a]);
      p = __dadd_rn(p, s2 * v2[a]);
    }
    double lo = p, hi = p;
    for (int off = 4; off >= 1; off >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (lane == 0) {
      aabb[(size_t)w * 6 + a] = lo;
      aabb[(size_t)w * 6 + 3 + a] = hi;
    }
  }
  // far-box test
  bool ok = !(sx < 0.01f || sy < 0.01f || sz < 0.01f || sx > 10.f || sy > 10.f || sz > 10.f);
  bool inside_any = false;
  if (ok) {
    // numpy: float32 size / 2. + python-float contact stays float32 (weak scalar), then promotes
    const float cf = (float)contact;
    const double ex = (double)__fadd_rn(sx * 0.5f, cf), ey = (double)__fadd_rn(sy * 0.5f, cf),
                 ez = (double)__fadd_rn(sz * 0.5f, cf);
    const float* hp = hip + (size_t)b * t * hip_stride;
    for (int f = lane; f < t; f += 32) {
      const double dx = (double)__ldg(hp + (size_t)f * hip_stride + 0) - cx;
      const double dy = (double)__ldg(hp + (size_t)f * hip_stride + 1) - cy;
      const double dz = (double)__ldg(hp + (size_t)f * hip_stride + 2) - cz;
      const double lx = dx * cs - dz * sn, lz = dx * sn + dz * cs;
      if (fabs(lx) <= ex && fabs(dy) <= ey && fabs(lz) <= ez) inside_any = true;
    }
  }
  inside_any = __any_sync(0xffffffffu, inside_any);
  if (lane == 0) nonempty[w] = (ok && inside_any) ? 1 : 0;
}

extern "C" int p2r_decode_boxes(const float* center, const float* log_size, const double* heading_sincos,
                                const float* hip, int hip_stride, int b, int k, int t, double contact,
                                double* corners, double* aabb, unsigned char* nonempty, void* stream) {
  P2R_CHECK_ARG(b >= 0 && k > 0 && t > 0 && hip_stride >= 3, "p2r_decode_boxes");
  if (b == 0) return 0;
  const int total = b * k;
  P2R_LAUNCH(decode_boxes_kernel, p2r_ceil_div((long long)total * 32, 128), 128, 0, (cudaStream_t)stream, k, t, center, log_size, heading_sincos, hip, hip_stride, contact, corners, aabb, nonempty, total);
  P2R_RETURN_LAUNCH("p2r_decode_boxes");
}

// ================================================================================================
// nms3d: greedy NMS over K <= 1024 axis-aligned boxes per scene (net_utils/nms.py:41-119).
// boxes (B,K,6) f64 [x1,y1,z1,x2,y2,z2], score (B,K) f64, valid (B,K) u8, cls (B,K) i32 or null
// -> keep (B,K) u8, order (B,K) i32 = picked indices in pick order (descending score), -1 padded.
// fp64 arithmetic in numpy's order, no contraction; suppression is strict  o > thr.
// Score ties: higher original index first (what np.argsort's result gives for the reference's
// short arrays); exact score ties are outside the parity contract.
// One CTA per scene: bitonic sort in smem, then a serial sweep where each step suppresses in parallel.
// ================================================================================================
#define NMS_MAXK 1024
__global__ void __launch_bounds__(256)
nms3d_kernel(int k, double thr, int old_type, const double* __restrict__ boxes, const double* __restrict__ score,
             const unsigned char* __restrict__ valid, const int* __restrict__ cls, unsigned char* __restrict__ keep,
             int* __restrict__ order) {
  __shared__ double s_score[NMS_MAXK];
  __shared__ int s_idx[NMS_MAXK];
  __shared__ unsigned char s_dead[NMS_MAXK];
  __shared__ int s_cur;
  const int b = blockIdx.x;
  const double* bx = boxes + (size_t)b * k * 6;
  int kp = 1;
  while (kp < k) kp <<= 1;
  for (int i = threadIdx.x; i < kp; i += blockDim.x) {
    const bool v = i < k && (valid == nullptr || valid[(size_t)b * k + i]);
    s_score[i] = v ? score[(size_t)b * k + i] : -INFINITY;
    s_idx[i] = v ? i : -1;
    s_dead[i] = v ? 0 : 1;
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    keep[(size_t)b * k + i] = 0;
    order[(size_t)b * k + i] = -1;
  }
  __syncthreads();
  // bitonic sort, descending by (score, idx)
  for (int size = 2; size <= kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < kp; i += blockDim.x) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const double si = s_score[i], sj = s_score[j];
          const int ii = s_idx[i], ij = s_idx[j];
          const bool i_before_j = (si > sj) || (si == sj && ii > ij);
          if (desc ? !i_before_j : i_before_j) {
            s_score[i] = sj; s_score[j] = si;
            s_idx[i] = ij; s_idx[j] = ii;
          }
        }
      }
      __syncthreads();
    }
  }
  // s_dead is indexed by sorted position from here on
  for (int i = threadIdx.x; i < kp; i += blockDim.x) s_dead[i] = s_idx[i] < 0 ? 1 : 0;
  __syncthreads();
  int npick = 0;
  for (int p = 0; p < k; ++p) {
    if (s_dead[p]) continue;  // uniform: all threads read the same smem byte after the barrier below
    const int i = s_idx[p];
    if (threadIdx.x == 0) {
      keep[(size_t)b * k + i] = 1;
      order[(size_t)b * k + npick] = i;
    }
    ++npick;
    const double x1 = bx[i * 6 + 0], y1 = bx[i * 6 + 1], z1 = bx[i * 6 + 2];
    const double x2 = bx[i * 6 + 3], y2 = bx[i * 6 + 4], z2 = bx[i * 6 + 5];
    const double area_i = __dmul_rn(__dmul_rn(__dsub_rn(x2, x1), __dsub_rn(y2, y1)), __dsub_rn(z2, z1));
    const int ci = cls ? cls[(size_t)b * k + i] : 0;
    for (int q = p + 1 + threadIdx.x; q < k; q += blockDim.x) {
      if (s_dead[q]) continue;
      const int j = s_idx[q];
      const double u1 = bx[j * 6 + 0], v1 = bx[j * 6 + 1], w1 = bx[j * 6 + 2];
      const double u2 = bx[j * 6 + 3], v2 = bx[j * 6 + 4], w2 = bx[j * 6 + 5];
      const double l = fmax(0.0, __dsub_rn(fmin(x2, u2), fmax(x1, u1)));
      const double w = fmax(0.0, __dsub_rn(fmin(y2, v2), fmax(y1, v1)));
      const double h = fmax(0.0, __dsub_rn(fmin(z2, w2), fmax(z1, w1)));
      const double inter = __dmul_rn(__dmul_rn(l, w), h);
      const double area_j = __dmul_rn(__dmul_rn(__dsub_rn(u2, u1), __dsub_rn(v2, v1)), __dsub_rn(w2, w1));
      double o = old_type ? __ddiv_rn(inter, area_j)
                          : __ddiv_rn(inter, __dsub_rn(__dadd_rn(area_i, area_j), inter));
      if (cls && cls[(size_t)b * k + j] != ci) o = __dmul_rn(o, 0.0);
      if (o > thr) s_dead[q] = 1;
    }
    __syncthreads();
  }
  (void)s_cur;
}

extern "C" int p2r_nms3d(const double* boxes, const double* score, const unsigned char* valid, const int* cls, int b,
                         int k, double thr, int old_type, unsigned char* keep, int* order, void* stream) {
  P2R_CHECK_ARG(b >= 0 && k > 0 && k <= NMS_MAXK, "p2r_nms3d");
  if (b == 0) return 0;
  P2R_LAUNCH(nms3d_kernel, b, 256, 0, (cudaStream_t)stream, k, thr, old_type, boxes, score, valid, cls, keep, order);
  P2R_RETURN_LAUNCH("p2r_nms3d");
}

// ================================================================================================
// box3d_iou: oriented-box IoU of every pair (box_util.py:90-118): Sutherland-Hodgman clip of the two
// XZ rectangles (same inside test and intersection formula as box_util.py:22-61), shoelace area,
// times the Y overlap.  c1 (P,8,3), c2 (G,8,3) f64, corner order of utils/tools.py:33-51
// -> iou3d (P,G), iou2d (P,G).  One thread per pair.
// ================================================================================================
__device__ __forceinline__ double poly_area(const double* px, const double* py, int cnt) {
  // 0.5*|dot(x, roll(y,1)) - dot(y, roll(x,1))|
  double a = 0.0, c = 0.0;
  for (int i = 0; i < cnt; ++i) {
    const int p = (i + cnt - 1) % cnt;
    a += px[i] * py[p];
    c += py[i] * px[p];
  }
  return 0.5 * fabs(a - c);
}

__global__ void __launch_bounds__(128)
box3d_iou_kernel(int np_, int ng, const double* __restrict__ c1, const double* __restrict__ c2,
                 double* __restrict__ iou3d, double* __restrict__ iou2d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= np_ * ng) return;
  const double* A = c1 + (size_t)(e / ng) * 24;
  const double* Bc = c2 + (size_t)(e % ng) * 24;
  const int perm[8] = {7, 6, 2, 3, 4, 5, 1, 0};
  // rect = permuted corners 3,2,1,0 -> original corners 3,2,6,7 (x,z)
  double sx[16], sy[16], tx[16], ty[16], qx[4], qy[4], rx[4], ry[4];
  for (int i = 0; i < 4; ++i) {
    const int o = perm[3 - i];
    rx[i] = A[o * 3 + 0]; ry[i] = A[o * 3 + 2];
    qx[i] = Bc[o * 3 + 0]; qy[i] = Bc[o * 3 + 2];
  }
  const double area1 = poly_area(rx, ry, 4), area2 = poly_area(qx, qy, 4);
  int cnt = 4;
  for (int i = 0; i < 4; ++i) { sx[i] = rx[i]; sy[i] = ry[i]; }
  double cp1x = qx[3], cp1y = qy[3];
  bool empty = false;
  for (int ce = 0; ce < 4 && !empty; ++ce) {
    const double cp2x = qx[ce], cp2y = qy[ce];
    int ocnt = 0;
    double s_x = sx[cnt - 1], s_y = sy[cnt - 1];
    for (int v = 0; v < cnt; ++v) {
      const double e_x = sx[v], e_y = sy[v];
      const bool in_e = (cp2x - cp1x) * (e_y - cp1y) > (cp2y - cp1y) * (e_x - cp1x);
      const bool in_s = (cp2x - cp1x) * (s_y - cp1y) > (cp2y - cp1y) * (s_x - cp1x);
      if (in_e != in_s) {
        const double dcx = cp1x - cp2x, dcy = cp1y - cp2y;
        const double dpx = s_x - e_x, dpy = s_y - e_y;
        const double n1 = cp1x * cp2y - cp1y * cp2x;
        const double n2 = s_x * e_y - s_y * e_x;
        const double n3 = 1.0 / (dcx * dpy - dcy * dpx);
        if (ocnt < 16) { tx[ocnt] = (n1 * dpx - n2 * dcx) * n3; ty[ocnt] = (n1 * dpy - n2 * dcy) * n3; ++ocnt; }
      }
      if (in_e && ocnt < 16) { tx[ocnt] = e_x; ty[ocnt] = e_y; ++ocnt; }
      s_x = e_x; s_y = e_y;
    }
    cnt = ocnt;
    for (int v = 0; v < cnt; ++v) { sx[v] = tx[v]; sy[v] = ty[v]; }
    cp1x = cp2x; cp1y = cp2y;
    if (cnt == 0) empty = true;
  }
  const double inter_area = (empty || cnt < 3) ? 0.0 : poly_area(sx, sy, cnt);
  const double i2 = inter_area / (area1 + area2 - inter_area);
  // permuted corner 0 = original 7 (top face), permuted 4 = original 4 (bottom face)
  const double ymax = fmin(A[7 * 3 + 1], Bc[7 * 3 + 1]);
  const double ymin = fmax(A[4 * 3 + 1], Bc[4 * 3 + 1]);
  const double inter_vol = inter_area * fmax(0.0, ymax - ymin);
  // box3d_vol on permuted corners: |p0-p1| * |p1-p2| * |p0-p4|  = |o7-o6| * |o6-o2| * |o7-o4|
  auto dist = [](const double* c, int i, int j) {
    const double dx = c[i * 3] - c[j * 3], dy = c[i * 3 + 1] - c[j * 3 + 1], dz = c[i * 3 + 2] - c[j * 3 + 2];
    return sqrt(dx * dx + dy * dy + dz * dz);
  };
  const double vol1 = dist(A, 7, 6) * dist(A, 6, 2) * dist(A, 7, 4);
  const double vol2 = dist(Bc, 7, 6) * dist(Bc, 6, 2) * dist(Bc, 7, 4);
  iou3d[e] = inter_vol / (vol1 + vol2 - inter_vol);
  iou2d[e] = i2;
}

extern "C" int p2r_box3d_iou(const double* corners1, const double* corners2, int np_, int ng, double* iou3d,
                             double* iou2d, void* stream) {
  P2R_CHECK_ARG(np_ >= 0 && ng >= 0, "p2r_box3d_iou");
  if (np_ == 0 || ng == 0) return 0;
  P2R_LAUNCH(box3d_iou_kernel, p2r_ceil_div((long long)np_ * ng, 128), 128, 0, (cudaStream_t)stream, np_, ng, corners1, corners2, iou3d, iou2d);
  P2R_RETURN_LAUNCH("p2r_box3d_iou");
}
