/* TEST HARNESS (CPU): runs the product's mixture-head arithmetic header (pose2room_b200/csrc/gmm_math.h) with the loop
 * structure of gmm_mix_fwd_kernel / gmm_mix_bwd_kernel (csrc/gmm_ops.cu; float64 storage, float32 logits), so
 * tests/test_gmm_math.py can hold it to torch autograd over the reference's formulas without a GPU.  Built by the test
 * with gcc; never shipped, never loaded by pose2room_b200/. */
#include <stddef.h>
#include "gmm_math.h"

void host_gmm_mix(const float* logits, const double* mu, const float* log_sigma, const double* eps, long long rows, int G,
                  int D, double* out) {
  double sig[P2RG_MAX_G * P2RG_MAX_D];
  for (int i = 0; i < G * D; ++i) sig[i] = (double)expf(log_sigma[i]);
  for (long long r = 0; r < rows; ++r) {
    double acc[P2RG_MAX_D] = {0.0, 0.0, 0.0, 0.0};
    for (int g = 0; g < G; ++g)
      p2rg_accumulate(p2rg_sigmoid(logits[r * G + g]), mu + g * D, sig + g * D, eps + (r * G + g) * D, D, acc);
    for (int c = 0; c < D; ++c) out[r * D + c] = acc[c];
  }
}

void host_gmm_mix_grad(const float* logits, const double* mu, const float* log_sigma, const double* eps, const double* dout,
                       long long rows, int G, int D, float* dlogits, double* dmu, float* dls) {
  double sig[P2RG_MAX_G * P2RG_MAX_D];
  for (int i = 0; i < G * D; ++i) sig[i] = (double)expf(log_sigma[i]);
  for (int g = 0; g < G; ++g) {
    double am[P2RG_MAX_D] = {0.0, 0.0, 0.0, 0.0}, al[P2RG_MAX_D] = {0.0, 0.0, 0.0, 0.0};
    for (long long r = 0; r < rows; ++r)
      dlogits[r * G + g] = (float)p2rg_backward(p2rg_sigmoid(logits[r * G + g]), mu + g * D, sig + g * D,
                                                 eps + (r * G + g) * D, dout + r * D, D, am, al);
    for (int c = 0; c < D; ++c) {
      dmu[g * D + c] = am[c];
      dls[g * D + c] = (float)al[c];
    }
  }
}
