// optim_ops.cu -- AdamW over many tensors in a few launches, capturable in a CUDA graph.
//
// The reference builds torch.optim.AdamW (models/optimizers.py:90); a whole-step CUDA graph needs its capturable form,
// whose fused kernel evaluates the two bias corrections (pow in double) per ELEMENT: 0.16 ms of a 7.1 ms step for 1.44 M
// parameters.  Same update rule here (torch/optim/adamw.py, amsgrad = False, maximize = False), the step count on the
// device, the corrections once per thread:
//     step' = step + 1
//     p    *= 1 - lr * weight_decay
//     m     = m + (g - m) * (1 - beta1)                (torch.lerp)
//     v     = beta2 * v + (1 - beta2) * g * g
//     p    -= (lr / (1 - beta1^step')) * m / (sqrt(v) / sqrt(1 - beta2^step') + eps)
// One launch covers up to P2R_ADAMW_CHUNK tensors (pointers travel as kernel parameters); blockIdx.y = tensor.
#include "p2r_common.cuh"

#define P2R_ADAMW_CHUNK 48
#define P2R_ADAMW_THREADS 256
#define P2R_ADAMW_PER_THREAD 4

struct AdamWChunk {
  float* p[P2R_ADAMW_CHUNK];
  const float* g[P2R_ADAMW_CHUNK];
  float* m[P2R_ADAMW_CHUNK];
  float* v[P2R_ADAMW_CHUNK];
  long long numel[P2R_ADAMW_CHUNK];   // negative: a float64 tensor of -numel elements (P2RNet has one: gmm_heading.mdn.mu)
};

__global__ void __launch_bounds__(P2R_ADAMW_THREADS)
adamw_chunk_kernel(const AdamWChunk c, const float* __restrict__ step, double lr, double beta1_d, double beta2_d, float eps,
                   double weight_decay) {
  const int t = blockIdx.y;
  const bool f64 = c.numel[t] < 0;
  const long long n = f64 ? -c.numel[t] : c.numel[t];
  const long long e0 = ((long long)blockIdx.x * P2R_ADAMW_THREADS + threadIdx.x) * P2R_ADAMW_PER_THREAD;
  if (e0 >= n) return;
  if (f64) {      // the same update in double, element by element
    const double s = (double)__ldg(step) + 1.0;
    const double bc2_sqrt = sqrt(1.0 - pow(beta2_d, s)), step_size = lr / (1.0 - pow(beta1_d, s));
    double* p = reinterpret_cast<double*>(c.p[t]);
    const double* g = reinterpret_cast<const double*>(c.g[t]);
    double* m = reinterpret_cast<double*>(c.m[t]);
    double* v = reinterpret_cast<double*>(c.v[t]);
    for (long long e = e0; e < n && e < e0 + P2R_ADAMW_PER_THREAD; ++e) {
      const double gg = g[e];
      const double mm = m[e] + (gg - m[e]) * (1.0 - beta1_d);
      const double vv = beta2_d * v[e] + (1.0 - beta2_d) * gg * gg;
      m[e] = mm;
      v[e] = vv;
      p[e] = p[e] * (1.0 - lr * weight_decay) - step_size * (mm / (sqrt(vv) / bc2_sqrt + (double)eps));
    }
    return;
  }
  // scalars in double like torch's Python side (1 - 0.999f is 1.3e-5 off 0.001), cast to float where torch casts
  const double s = (double)__ldg(step) + 1.0;
  const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2_d, s));
  const float step_size = (float)(lr / (1.0 - pow(beta1_d, s))), decay = (float)(1.0 - lr * weight_decay);
  const float beta2 = (float)beta2_d, omb1 = (float)(1.0 - beta1_d), omb2 = (float)(1.0 - beta2_d);
  float* p = c.p[t];
  const float* g = c.g[t];
  float* m = c.m[t];
  float* v = c.v[t];
  const bool vec = e0 + P2R_ADAMW_PER_THREAD <= n &&
                   (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  float pp[4], gg[4], mm[4], vv[4];
  if (vec) {
    const float4 a = *reinterpret_cast<const float4*>(p + e0), b = __ldg(reinterpret_cast<const float4*>(g + e0));
    const float4 cm = *reinterpret_cast<const float4*>(m + e0), cv = *reinterpret_cast<const float4*>(v + e0);
    pp[0] = a.x; pp[1] = a.y; pp[2] = a.z; pp[3] = a.w;
    gg[0] = b.x; gg[1] = b.y; gg[2] = b.z; gg[3] = b.w;
    mm[0] = cm.x; mm[1] = cm.y; mm[2] = cm.z; mm[3] = cm.w;
    vv[0] = cv.x; vv[1] = cv.y; vv[2] = cv.z; vv[3] = cv.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = e0 + i < n;
      pp[i] = ok ? p[e0 + i] : 0.f;
      gg[i] = ok ? g[e0 + i] : 0.f;
      mm[i] = ok ? m[e0 + i] : 0.f;
      vv[i] = ok ? v[e0 + i] : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pp[i] *= decay;
    mm[i] = mm[i] + (gg[i] - mm[i]) * omb1;
    vv[i] = beta2 * vv[i] + omb2 * gg[i] * gg[i];
    const float denom = sqrtf(vv[i]) / bc2_sqrt + eps;
    pp[i] -= step_size * (mm[i] / denom);
  }
  if (vec) {
    *reinterpret_cast<float4*>(p + e0) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + e0) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + e0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (e0 + i < n) { p[e0 + i] = pp[i]; m[e0 + i] = mm[i]; v[e0 + i] = vv[i]; }
  }
}

__global__ void adamw_step_inc_kernel(float* step) { *step += 1.f; }

// params / grads / exp_avg / exp_avg_sq: HOST arrays of n device pointers (dense), numel: host array of n counts -- positive for
// float32 tensors, NEGATIVE for float64 ones (parameter, gradient and both moments in double).
// step: DEVICE float, the number of updates already applied; incremented by 1 after all tensors are updated.
extern "C" int p2r_adamw_step(int n, const void* const* params, const void* const* grads, void* const* exp_avg,
                              void* const* exp_avg_sq, const long long* numel, float* step, double lr, double beta1,
                              double beta2, double eps, double weight_decay, void* stream) {
  P2R_CHECK_ARG(n >= 0 && step != nullptr && (n == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), "p2r_adamw_step");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i0 = 0; i0 < n; i0 += P2R_ADAMW_CHUNK) {
    AdamWChunk c;
    const int cnt = n - i0 < P2R_ADAMW_CHUNK ? n - i0 : P2R_ADAMW_CHUNK;
    long long mx = 0;
    for (int i = 0; i < P2R_ADAMW_CHUNK; ++i) {
      const bool live = i < cnt;
      c.p[i] = live ? (float*)params[i0 + i] : nullptr;
      c.g[i] = live ? (const float*)grads[i0 + i] : nullptr;
      c.m[i] = live ? (float*)exp_avg[i0 + i] : nullptr;
      c.v[i] = live ? (float*)exp_avg_sq[i0 + i] : nullptr;
      c.numel[i] = live ? numel[i0 + i] : 0;
      const long long an = c.numel[i] < 0 ? -c.numel[i] : c.numel[i];
      if (an > mx) mx = an;
    }
    if (mx == 0) continue;
    const long long per_block = (long long)P2R_ADAMW_THREADS * P2R_ADAMW_PER_THREAD;
    dim3 grid((unsigned)((mx + per_block - 1) / per_block), (unsigned)cnt);
    adamw_chunk_kernel<<<grid, P2R_ADAMW_THREADS, 0, st>>>(c, step, lr, beta1, beta2, (float)eps, weight_decay);
  }
  adamw_step_inc_kernel<<<1, 1, 0, st>>>(step);
  P2R_RETURN_LAUNCH("p2r_adamw_step");
}
