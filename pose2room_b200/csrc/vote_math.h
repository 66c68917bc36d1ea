/*
 * vote_math.h -- per-row arithmetic of the fused vote-tail kernels (vote_ops.cu).
 *
 * Restates the end of CenterVoteModule.forward (ref: models/p2rnet/modules/vote_center.py:52-58) and the L2
 * normalisation P2RNet applies to the vote features right after it (ref: models/p2rnet/modules/network.py:89-90):
 *     vote_xyz  = seed_xyz + net[:, 0:3]
 *     v         = seed_features + net[:, 3:]
 *     vote_feat = v / ||v||_2
 * and the gradient:  dv = (g - vote_feat * <vote_feat, g>) / ||v||,  d net = [g_xyz, dv],  d seed_features = dv.
 * float32 like the reference; the two row reductions are the only place the summation order differs from torch's.
 * Compiles for the device (nvcc) and the host (tests).
 */
#ifndef P2R_VOTE_MATH_H
#define P2R_VOTE_MATH_H

#ifdef __CUDACC__
#define P2RV_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define P2RV_HD static inline
#endif

/* forward of one element once the row's squared norm is known */
P2RV_HD float p2rv_normalise(float v, float norm) { return v / norm; }
/* backward of one element: g = upstream, y = normalised feature, dot = <y, g> of the row */
P2RV_HD float p2rv_dnormalise(float g, float y, float dot, float norm) { return (g - y * dot) / norm; }

#endif /* P2R_VOTE_MATH_H */
