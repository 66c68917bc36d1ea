/*
 * gmm_math.h -- per-element arithmetic of the fused Gaussian-mixture head kernels (gmm_ops.cu).
 *
 * Restates the training-time point prediction of MixtureDensityHead (ref: models/p2rnet/modules/mdn.py:36-84 with
 * n_samples = 1, sample_pi = False, central_tendency = 'mean', the configuration proposal_net.py:141-147 builds):
 *     pi      = sigmoid(logit)                                   mdn.py:31-34
 *     sample  = mu + exp(log_sigma) * eps,  eps ~ N(0,1)         mdn.py:38-46   (eps is drawn by torch, not here)
 *     out[r]  = sum_g pi[r,g] * sample[r,g,:]                    mdn.py:64-66, :80
 * and its gradient.  Evaluated in float64 whatever the storage types (the heading head IS float64 in the reference: its
 * mu grid is a float64 array, proposal_net.py:131; the centre / size heads are float32 there -- agreement ~1e-7).
 *
 * The same source compiles for the device (nvcc) and for the host (gcc; tests/test_gmm_math.py).
 */
#ifndef P2R_GMM_MATH_H
#define P2R_GMM_MATH_H

#ifdef __CUDACC__
#define P2RG_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define P2RG_HD static inline
#endif

#define P2RG_MAX_D 4      /* out_dim of a head (3 for centre / size, 2 for heading) */
#define P2RG_MAX_G 256    /* mixture components per head (100 in the reference YAML) */

P2RG_HD double p2rg_sigmoid(float logit) { return 1.0 / (1.0 + exp(-(double)logit)); }

/* forward contribution of component g to row r: acc[d] += pi * (mu[d] + sigma[d] * eps[d]) */
P2RG_HD void p2rg_accumulate(double pi, const double* mu, const double* sigma, const double* eps, int d, double* acc) {
  for (int c = 0; c < d; ++c) acc[c] += pi * (mu[c] + sigma[c] * eps[c]);
}

/* backward of one (r, g): returns d loss / d logit[r,g]; adds this row's share to dmu[d] / dls[d] of component g */
P2RG_HD double p2rg_backward(double pi, const double* mu, const double* sigma, const double* eps, const double* dout,
                             int d, double* dmu, double* dls) {
  double dpi = 0.0;
  for (int c = 0; c < d; ++c) {
    dpi += dout[c] * (mu[c] + sigma[c] * eps[c]);
    dmu[c] += dout[c] * pi;
    dls[c] += dout[c] * pi * eps[c] * sigma[c];
  }
  return dpi * pi * (1.0 - pi);
}

#endif /* P2R_GMM_MATH_H */
