// loss_ops.cu -- BoxNetDetectionLoss (ref: models/loss.py:42-189) as ONE forward launch and ONE backward launch.
//
// The reference evaluates the detection loss as ~100 tiny torch kernels forward and ~150 backward (three nn_distance
// calls with materialised (B,N,M,3) tiles, gathers, masks, two cross entropies, ten reductions); inside the captured
// train step that is a chain of launch latencies, not work: 16 384 seeds x 75 distances + 4096 proposals x 10 boxes.
// Here: blocks [0, B) own the proposals of one sample (ground truth staged in shared memory, the per-ground-truth
// nearest-proposal search as a block-wide lexicographic arg-min), blocks [B, B + ceil(B*S/128)) own 128 seeds each.
// Every block leaves its P2RL_NSUM float64 partial sums in global memory; the last block to finish (threadfence +
// counter) adds them in block order -- a deterministic reduction -- and writes the ten reported numbers and the four
// reciprocals the backward needs.  The forward also writes the un-normalised gradient of every sum, so the backward is
// one elementwise pass.  Arithmetic: loss_math.h (shared with the host-compiled CPU test).
//
// Opt-in (P2R_FUSED_LOSS=1 in pose2room_b200/p2rnet/loss.py) until it has been A/B-ed on a B200.
#include "p2r_common.cuh"
#include "p2r_b200.h"
#include "loss_math.h"

#define P2RL_THREADS 128

struct DetLossArgs {
  // predictions
  const float* vote_xyz;        // [B,S,3]
  const float* center;          // [B,P,3]
  const float* size;            // [B,P,3]
  const void* heading;          // [B,P,2] f64 or f32
  const float* obj;             // [B,P,2], row stride obj_stride
  const float* sem;             // [B,P,C], row stride sem_stride
  const float* agg;             // [B,P,3]
  const float* skeleton;        // [B,S,J,3]
  const long long* seed_inds;   // [B,S]
  // ground truth
  const float* vote_label;      // [B,T,J,9]
  const long long* vote_mask;   // [B,T,J]
  const float* gt_center;       // [B,G,3]
  const float* gt_mask;         // [B,G]
  const float* gt_size;         // [B,G,3]
  const float* gt_heading;      // [B,G,2]
  const long long* gt_cls;      // [B,G]
  int B, S, J, T, P, G, C, origin, heading_f64, obj_stride, sem_stride;
  // outputs
  float* out32;                 // [P2RL_NOUT32]
  double* out64;                // [P2RL_NOUT64]
  double* scales;               // [P2RL_NSCALE]
  float* u_vote;                // [B,S,3]
  float* u_c1;                  // [B,P,3]
  float* u_c2;                  // [B,P,3]
  float* u_size;                // [B,P,3]
  void* u_head;                 // [B,P,2] like heading
  float* u_obj;                 // [B,P,2]
  float* u_sem;                 // [B,P,C]
  double* partials;             // [gridDim.x, P2RL_NSUM] scratch
  unsigned int* counter;        // zero on entry
};

__device__ __forceinline__ double p2rl_warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  return v;
}

__global__ void __launch_bounds__(P2RL_THREADS) detection_loss_kernel(DetLossArgs a) {
  P2R_DYN_SMEM(float, s_center);                                       // [P*3], proposal blocks only
  __shared__ float s_gc[P2RL_MAX_GT * 3], s_gm[P2RL_MAX_GT], s_gs[P2RL_MAX_GT * 3], s_gh[P2RL_MAX_GT * 2];
  __shared__ long long s_gcls[P2RL_MAX_GT];
  __shared__ float s_d2[P2RL_MAX_GT];
  __shared__ int s_i2[P2RL_MAX_GT];
  __shared__ float s_wb[P2RL_THREADS / 32];
  __shared__ int s_wi[P2RL_THREADS / 32];
  __shared__ double s_red[P2RL_THREADS / 32][P2RL_NSUM];
  __shared__ bool s_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double sums[P2RL_NSUM];
#pragma unroll
  for (int k = 0; k < P2RL_NSUM; ++k) sums[k] = 0.0;

  if ((int)blockIdx.x < a.B) {
    // ---------------------------------------------------------------- the proposals of sample b
    const int b = blockIdx.x, P = a.P, G = a.G, C = a.C;
    for (int i = tid; i < G * 3; i += P2RL_THREADS) {
      s_gc[i] = __ldg(a.gt_center + (size_t)b * G * 3 + i);
      s_gs[i] = __ldg(a.gt_size + (size_t)b * G * 3 + i);
    }
    for (int i = tid; i < G * 2; i += P2RL_THREADS) s_gh[i] = __ldg(a.gt_heading + (size_t)b * G * 2 + i);
    for (int i = tid; i < G; i += P2RL_THREADS) {
      s_gm[i] = __ldg(a.gt_mask + (size_t)b * G + i);
      s_gcls[i] = __ldg(a.gt_cls + (size_t)b * G + i);
    }
    for (int i = tid; i < P * 3; i += P2RL_THREADS) s_center[i] = __ldg(a.center + (size_t)b * P * 3 + i);
    __syncthreads();

    for (int p = tid; p < P; p += P2RL_THREADS) {
      const size_t r = (size_t)b * P + p;
      const void* hd = a.heading_f64 ? (const void*)((const double*)a.heading + r * 2)
                                     : (const void*)((const float*)a.heading + r * 2);
      void* uh = a.heading_f64 ? (void*)((double*)a.u_head + r * 2) : (void*)((float*)a.u_head + r * 2);
      p2rl_proposal(a.agg + r * 3, s_center + p * 3, a.size + r * 3, hd, a.heading_f64, a.obj + r * a.obj_stride,
                    a.sem + r * a.sem_stride, C, s_gc, s_gm, s_gs, s_gh, s_gcls, G, a.u_c1 + r * 3, a.u_size + r * 3, uh,
                    a.u_obj + r * 2, a.u_sem + r * C, sums);
    }
    // ground-truth side of the centre loss: for every slot g the nearest proposal, first minimum (torch.min)
    for (int g = 0; g < G; ++g) {
      float best = __int_as_float(0x7f800000);
      int bi = 0x7fffffff;
      for (int p = tid; p < P; p += P2RL_THREADS) {
        const float d = p2rl_sqdist3(s_center + p * 3, s_gc + g * 3);
        if (bi == 0x7fffffff || d < best) { best = d; bi = p; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ob < best || (ob == best && oi < bi))) { best = ob; bi = oi; }
      }
      if (lane == 0) { s_wb[warp] = best; s_wi[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        float fb = s_wb[0];
        int fi = s_wi[0];
        for (int w = 1; w < P2RL_THREADS / 32; ++w) {
          const float ob = s_wb[w];
          const int oi = s_wi[w];
          if (oi != 0x7fffffff && (fi == 0x7fffffff || ob < fb || (ob == fb && oi < fi))) { fb = ob; fi = oi; }
        }
        s_d2[g] = fb;
        s_i2[g] = fi;
        p2rl_gt_side(fb, s_gm[g], sums);
      }
      __syncthreads();
    }
    for (int p = tid; p < P; p += P2RL_THREADS) {
      float u[3] = {0.f, 0.f, 0.f};
      for (int g = 0; g < G; ++g)
        if (s_i2[g] == p) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            u[c] = P2RL_FADD(u[c], s_gm[g] * 2.0f * P2RL_FSUB(s_center[p * 3 + c], s_gc[g * 3 + c]));
        }
      const size_t r = (size_t)b * P + p;
#pragma unroll
      for (int c = 0; c < 3; ++c) a.u_c2[r * 3 + c] = u[c];
    }
  } else {
    // ---------------------------------------------------------------- 128 seeds
    const long long seed = (long long)(blockIdx.x - a.B) * P2RL_THREADS + tid;
    if (seed < (long long)a.B * a.S) {
      const int b = (int)(seed / a.S);
      const long long fr = __ldg(a.seed_inds + seed);
      const size_t row = ((size_t)b * a.T + (size_t)fr) * a.J + a.origin;
      float gv[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) gv[i] = __ldg(a.vote_label + row * 9 + i);
      p2rl_seed(a.skeleton + (size_t)seed * a.J * 3, a.J, a.origin, gv, __ldg(a.vote_mask + row),
                a.vote_xyz + (size_t)seed * 3, a.u_vote + (size_t)seed * 3, sums);
    }
  }

  // ---- block partial sums -> global; the last block reduces them in block order and finalises -----------------
#pragma unroll
  for (int k = 0; k < P2RL_NSUM; ++k) {
    const double v = p2rl_warp_sum(sums[k]);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  if (tid < P2RL_NSUM) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < P2RL_THREADS / 32; ++w) t += s_red[w][tid];
    a.partials[(size_t)blockIdx.x * P2RL_NSUM + tid] = t;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* s_tot = &s_red[0][0];
  if (tid < P2RL_NSUM) {
    double t = 0.0;
    for (unsigned blk = 0; blk < gridDim.x; ++blk) t += __ldcg(a.partials + (size_t)blk * P2RL_NSUM + tid);
    s_tot[tid] = t;
  }
  __syncthreads();
  if (tid == 0) p2rl_finalize(s_tot, (double)a.B * (double)a.P, a.out32, a.out64, a.scales);
}

extern "C" int p2r_detection_loss(const float* vote_xyz, const float* center, const float* size, const void* heading,
                                  int heading_f64, const float* obj, int obj_stride, const float* sem, int sem_stride,
                                  const float* agg, const float* skeleton, const long long* seed_inds,
                                  const float* vote_label, const long long* vote_mask, const float* gt_center,
                                  const float* gt_mask, const float* gt_size, const float* gt_heading,
                                  const long long* gt_cls, int b, int s, int j, int t, int p, int g, int c, int origin,
                                  float* out32, double* out64, double* scales, float* u_vote, float* u_c1, float* u_c2,
                                  float* u_size, void* u_head, float* u_obj, float* u_sem, double* workspace,
                                  long long workspace_doubles, void* stream) {
  P2R_CHECK_ARG(b >= 0 && s > 0 && j > 0 && t > 0 && p > 0 && g > 0 && c > 0, "p2r_detection_loss");
  P2R_CHECK_ARG(g <= P2RL_MAX_GT && origin >= 0 && origin < j, "p2r_detection_loss");
  P2R_CHECK_ARG(obj_stride >= 2 && sem_stride >= c, "p2r_detection_loss");
  P2R_CHECK_ARG((size_t)p * 3 * sizeof(float) <= 160 * 1024, "p2r_detection_loss");
  if (b == 0) return 0;
  const int grid = b + p2r_ceil_div((long long)b * s, P2RL_THREADS);
  P2R_CHECK_ARG(workspace_doubles >= p2r_detection_loss_workspace(b, s), "p2r_detection_loss (workspace too small)");
  DetLossArgs a;
  a.vote_xyz = vote_xyz; a.center = center; a.size = size; a.heading = heading; a.obj = obj; a.sem = sem; a.agg = agg;
  a.skeleton = skeleton; a.seed_inds = seed_inds; a.vote_label = vote_label; a.vote_mask = vote_mask;
  a.gt_center = gt_center; a.gt_mask = gt_mask; a.gt_size = gt_size; a.gt_heading = gt_heading; a.gt_cls = gt_cls;
  a.B = b; a.S = s; a.J = j; a.T = t; a.P = p; a.G = g; a.C = c; a.origin = origin; a.heading_f64 = heading_f64;
  a.obj_stride = obj_stride; a.sem_stride = sem_stride;
  a.out32 = out32; a.out64 = out64; a.scales = scales; a.u_vote = u_vote; a.u_c1 = u_c1; a.u_c2 = u_c2;
  a.u_size = u_size; a.u_head = u_head; a.u_obj = u_obj; a.u_sem = u_sem;
  a.counter = reinterpret_cast<unsigned int*>(workspace);          // first 8 bytes: the block counter (zero on entry)
  a.partials = workspace + 1;
  const size_t smem = (size_t)p * 3 * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(detection_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { p2r_set_last_error("p2r_detection_loss", (int)e); return (int)e; }
  }
  P2R_LAUNCH(detection_loss_kernel, grid, P2RL_THREADS, smem, (cudaStream_t)stream, a);
  P2R_RETURN_LAUNCH("p2r_detection_loss");
}

extern "C" long long p2r_detection_loss_workspace(int b, int s) {
  const long long grid = (long long)b + ((long long)b * s + P2RL_THREADS - 1) / P2RL_THREADS;
  return 1 + grid * P2RL_NSUM;
}

// ---------------------------------------------------------------------------------------------------------------
// backward: grad = upstream(term) * scale * u, all six differentiable inputs in one launch.
struct DetLossGradArgs {
  const float* g32;      // [P2RL_NOUT32] upstream gradients of the float32 outputs
  const double* g64;     // [P2RL_NOUT64]
  const double* scales;  // [P2RL_NSCALE]
  const float *u_vote, *u_c1, *u_c2, *u_size, *u_obj, *u_sem;
  const void* u_head;
  float *d_vote, *d_center, *d_size, *d_obj, *d_sem;
  void* d_head;
  long long n_vote, n_p3, n_p2, n_sem;     // element counts: B*S*3, B*P*3, B*P*2, B*P*C
  int heading_f64;
};

__global__ void __launch_bounds__(256) detection_loss_grad_kernel(DetLossGradArgs a) {
  __shared__ double s_w[P2RL_NTERM];
  if (threadIdx.x == 0) p2rl_term_weights(a.g32, a.g64, s_w);
  __syncthreads();
  const double sv = a.scales[P2RL_SC_VOTE], so = a.scales[P2RL_SC_OBJMASK], sp = a.scales[P2RL_SC_POS],
               sb = a.scales[P2RL_SC_BOXMASK];
  const long long total = a.n_vote + 2 * a.n_p3 + 2 * a.n_p2 + a.n_sem;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    long long i = e;
    if (i < a.n_vote) { a.d_vote[i] = (float)(s_w[P2RL_T_VOTE] * sv * (double)a.u_vote[i]); continue; }
    i -= a.n_vote;
    if (i < a.n_p3) {
      a.d_center[i] = (float)(s_w[P2RL_T_CENTER] * 0.5 * (sp * (double)a.u_c1[i] + sb * (double)a.u_c2[i]));
      continue;
    }
    i -= a.n_p3;
    if (i < a.n_p3) { a.d_size[i] = (float)(s_w[P2RL_T_SIZE] * sp * (double)a.u_size[i]); continue; }
    i -= a.n_p3;
    if (i < a.n_p2) {
      if (a.heading_f64) ((double*)a.d_head)[i] = s_w[P2RL_T_HEADING] * sp * ((const double*)a.u_head)[i];
      else ((float*)a.d_head)[i] = (float)(s_w[P2RL_T_HEADING] * sp * (double)((const float*)a.u_head)[i]);
      continue;
    }
    i -= a.n_p2;
    if (i < a.n_p2) { a.d_obj[i] = (float)(s_w[P2RL_T_OBJ] * so * (double)a.u_obj[i]); continue; }
    i -= a.n_p2;
    a.d_sem[i] = (float)(s_w[P2RL_T_SEM] * sp * (double)a.u_sem[i]);
  }
}

extern "C" int p2r_detection_loss_grad(const float* g32, const double* g64, const double* scales, const float* u_vote,
                                       const float* u_c1, const float* u_c2, const float* u_size, const void* u_head,
                                       int heading_f64, const float* u_obj, const float* u_sem, int b, int s, int p,
                                       int c, float* d_vote, float* d_center, float* d_size, void* d_head,
                                       float* d_obj, float* d_sem, void* stream) {
  P2R_CHECK_ARG(b >= 0 && s > 0 && p > 0 && c > 0, "p2r_detection_loss_grad");
  if (b == 0) return 0;
  DetLossGradArgs a;
  a.g32 = g32; a.g64 = g64; a.scales = scales; a.u_vote = u_vote; a.u_c1 = u_c1; a.u_c2 = u_c2; a.u_size = u_size;
  a.u_obj = u_obj; a.u_sem = u_sem; a.u_head = u_head; a.d_vote = d_vote; a.d_center = d_center; a.d_size = d_size;
  a.d_obj = d_obj; a.d_sem = d_sem; a.d_head = d_head; a.heading_f64 = heading_f64;
  a.n_vote = (long long)b * s * 3; a.n_p3 = (long long)b * p * 3; a.n_p2 = (long long)b * p * 2;
  a.n_sem = (long long)b * p * c;
  const long long total = a.n_vote + 2 * a.n_p3 + 2 * a.n_p2 + a.n_sem;
  int grid = p2r_ceil_div(total, 256);
  if (grid > P2R_SM_COUNT * 8) grid = P2R_SM_COUNT * 8;
  P2R_LAUNCH(detection_loss_grad_kernel, grid, 256, 0, (cudaStream_t)stream, a);
  P2R_RETURN_LAUNCH("p2r_detection_loss_grad");
}
