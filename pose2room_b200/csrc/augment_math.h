/*
 * augment_math.h -- per-joint arithmetic of the sample -> batch kernel (dataloader_ops.cu).
 *
 * Restates, with every rounding step pinned, what the reference does per raw sample in numpy
 * (ref: models/p2rnet/dataloader.py:31-84 augment_data, :128-131 frame picking).  The reference mixes
 * float32 storage with float64 arithmetic and the place where a value is rounded to float32 depends on
 * the flip flag (see the dtype notes in oracle/dataloader_ref.py); the order below reproduces numpy's
 * results bit for bit:
 *   - a 3-term dot product is  fma(x2,m2, fma(x1,m1, x0*m0))  in float64 -- measured: np.dot on this
 *     image's OpenBLAS 0.3.30 accumulates sequentially with FMA (the other five orders differ);
 *   - the rotated vote END POINT is rounded to float32 before the rotated joint (float64) is subtracted
 *     (`point_votes_end = np.zeros_like(votes)` is a float32 array, dataloader.py:64);
 *   - joint + vote is a float32 addition for an un-flipped sample and a float64 addition for a flipped one.
 *
 * The same source compiles for the device (nvcc) and for the host (gcc -ffp-contract=off, used only by
 * tests/test_dataloader_math.py to check this arithmetic against the reference goldens without a GPU).
 */
#ifndef P2R_AUGMENT_MATH_H
#define P2R_AUGMENT_MATH_H

#ifdef __CUDACC__
#define P2R_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define P2R_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define P2R_DMUL(a, b) __dmul_rn((a), (b))
#define P2R_DADD(a, b) __dadd_rn((a), (b))
#define P2R_DSUB(a, b) __dsub_rn((a), (b))
#define P2R_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define P2R_FADD(a, b) __fadd_rn((a), (b))
#define P2R_FSUB(a, b) __fsub_rn((a), (b))
#define P2R_D2F(a) __double2float_rn(a)
#else
#define P2R_DMUL(a, b) ((a) * (b))
#define P2R_DADD(a, b) ((a) + (b))
#define P2R_DSUB(a, b) ((a) - (b))
#define P2R_DFMA(a, b, c) fma((a), (b), (c))
#define P2R_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define P2R_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define P2R_D2F(a) ((float)(a))
#endif

/* Per-batch-item parameter block, 16 doubles, built on the host (pose2room_b200/dataloader.py). */
#define P2R_AUG_STRIDE 16
#define P2R_AUG_ENABLED 0 /* 0.0 = val/test: raw values are copied through unchanged            */
#define P2R_AUG_FLIP 1    /* 1.0 = x<->z flip (dataloader.py:43-52)                               */
#define P2R_AUG_ROT 2     /* 9 doubles, row-major rot_func(theta) (dataloader.py:26-28)           */
#define P2R_AUG_SHIFT 11  /* 3 doubles, offset_func(scale) = (scale, 0*scale, scale) (:29,:78)    */
#define P2R_AUG_FLOOR 14  /* floor height for the optional 4th channel (dataloader.py:112-115)    */

/* Raw frame that becomes network frame t: np.linspace(0, n_raw-1, num_frames).round().astype(uint16)
 * (dataloader.py:128).  linspace = arange(num)*step + 0 with step = (n_raw-1)/(num-1), last element
 * forced to the end point; round = half-to-even; the uint16 cast wraps like the reference.          */
P2R_HD int p2r_frame_id(int n_raw, int num_frames, int t) {
  if (num_frames <= 1) return 0;
  const double stop = (double)(n_raw - 1);
  const double step = stop / (double)(num_frames - 1);
  const double y = (t == num_frames - 1) ? stop : P2R_DMUL((double)t, step);
  return (int)(((long long)rint(y)) & 0xFFFF);
}

/* out[k] = sum_i x[i] * M[i][k], numpy's accumulation order. */
P2R_HD void p2r_dot3(const double x[3], const double* M, double out[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double acc = P2R_DMUL(x[0], M[0 * 3 + k]);
    acc = P2R_DFMA(x[1], M[1 * 3 + k], acc);
    acc = P2R_DFMA(x[2], M[2 * 3 + k], acc);
    out[k] = acc;
  }
}

/* One joint of one frame.  jr: 3 raw floats; vr: 10 raw floats (mask + 3 votes); p: parameter block.
 * oj: out_c (3 or 4) floats; ov: 9 floats; returns the vote mask (float -> int64 truncation).       */
P2R_HD long long p2r_augment_joint(const float* jr, const float* vr, const double* p, int out_c, float* oj,
                                   float* ov) {
  const long long mask = (long long)vr[0];
  if (p[P2R_AUG_ENABLED] == 0.0) {
    oj[0] = jr[0]; oj[1] = jr[1]; oj[2] = jr[2];
    /* un-augmented joints stay float32, and so does np.percentile's result: float32 subtraction */
    if (out_c == 4) oj[3] = P2R_FSUB(jr[1], P2R_D2F(p[P2R_AUG_FLOOR]));
#pragma unroll
    for (int i = 0; i < 9; ++i) ov[i] = vr[1 + i];
    return mask;
  }
  const double FLIP[9] = {0.0, 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
  const int flip = p[P2R_AUG_FLIP] != 0.0;
  const double* R = p + P2R_AUG_ROT;
  double j[3] = {(double)jr[0], (double)jr[1], (double)jr[2]};
  float v[9];
  if (flip) {
    double t[3];
    p2r_dot3(j, FLIP, t);
    j[0] = t[0]; j[1] = t[1]; j[2] = t[2];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const double x[3] = {(double)vr[1 + 3 * s], (double)vr[2 + 3 * s], (double)vr[3 + 3 * s]};
      p2r_dot3(x, FLIP, t);
      v[3 * s] = P2R_D2F(t[0]); v[3 * s + 1] = P2R_D2F(t[1]); v[3 * s + 2] = P2R_D2F(t[2]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = vr[1 + i];
  }
  double jrot[3];
  p2r_dot3(j, R, jrot);
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    double e[3], erot[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
      e[i] = flip ? P2R_DADD(j[i], (double)v[3 * s + i]) : (double)P2R_FADD(jr[i], v[3 * s + i]);
    p2r_dot3(e, R, erot);
#pragma unroll
    for (int i = 0; i < 3; ++i) ov[3 * s + i] = P2R_D2F(P2R_DSUB((double)P2R_D2F(erot[i]), jrot[i]));
  }
  double moved[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    moved[i] = P2R_DADD(jrot[i], p[P2R_AUG_SHIFT + i]);
    oj[i] = P2R_D2F(moved[i]);
  }
  if (out_c == 4) oj[3] = P2R_D2F(P2R_DSUB(moved[1], p[P2R_AUG_FLOOR]));
  return mask;
}

#endif /* P2R_AUGMENT_MATH_H */
