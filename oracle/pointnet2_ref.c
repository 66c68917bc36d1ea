/*
 * oracle/pointnet2_ref.c -- CPU restatement of the reference's nine native PointNet++ ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under pose2room_b200/ links, loads or calls this file;
 * it is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * The reference has NO CPU path for these ops (every entry point asserts "CPU not supported",
 * _ext-src/src/sampling.cpp:34,61,83, ball_query.cpp:28, group_points.cpp:32,58,
 * interpolate.cpp:36,66,95), so this file restates the CUDA kernels' arithmetic, including
 * their quirks, in scalar C.  File:line citations are into
 * /root/reference/external/pointnet2_ops_lib/pointnet2_ops/_ext-src/.
 *
 * Floating point: the reference is compiled by nvcc with default -fmad=true; the SASS of the
 * reference built for sm_100a (oracle/build_ref_ext.py; cuobjdump -sass) evaluates
 *     (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f)          [x, y, z terms in source order]
 * as  FMUL t=dy*dy ; FFMA t=dx*dx+t ; FFMA t=dz*dz+t   (the compiler fuses the FIRST product of each
 * addition into the FMA, so the SECOND product of the first addition is the plain multiply; checked on
 * the loads' address offsets in ball_query, three_nn and FPS; three_interpolate likewise evaluates
 * p1*w1 + p2*w2 + p3*w3 as t=p2*w2 ; t=fma(p1,w1,t) ; t=fma(p3,w3,t)).  sqdist3() reproduces exactly that
 * with fmaf(), so index decisions (strict < and > compares) AND the returned distances agree bit for bit.
 *
 * Pinning: the reference's own tests hold no golden vectors for these ops;
 * the pin is the reference kernels themselves, run on the GPU box from oracle/_ref/ and compared
 * with this file in tests/test_ref_ext_gpu.py.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/libp2r_oracle.so oracle/pointnet2_ref.c -lm
 *        (-ffp-contract=off: only the explicit fmaf() calls below may fuse).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  float t = dy * dy;
  t = fmaf(dx, dx, t);
  t = fmaf(dz, dz, t);
  return t;
}

/* include/cuda_utils.h:15-19 -- largest power of two <= work_size, clamped to [1, 512]. */
int p2r_ref_opt_n_threads(int work_size) {
  int pow_2 = (int)(log((double)work_size) / log(2.0));
  int t = 1 << pow_2;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

/* sampling_gpu.cu:70-173 furthest_point_sampling_kernel<block_size>; host side sampling.cpp:66-87
 * (idxs zero-initialised, temp filled with 1e10).
 *  - starts at index 0 (:88-89);
 *  - points with x^2+y^2+z^2 <= 1e-3 are skipped: never selectable, temp untouched (:100-101);
 *  - per-thread running arg-max over k = tid, tid+bs, ... with strict > (:108-109), best=-1, besti=0;
 *  - block tree reduction (__update :59-65) keeps slot t over slot t+off on equal values.  Two equal
 *    maxima therefore meet at the level of the LOWEST bit in which their thread ids differ and the
 *    id with a 0 there survives: ties go to the smallest BIT-REVERSED (k mod bs), then the smallest k
 *    (not simply the lowest thread id).  The tree is simulated literally below.
 */
void p2r_ref_furthest_point_sampling(int b, int n, int m, const float *dataset, int32_t *idxs) {
  if (m <= 0) return;
  int bs = p2r_ref_opt_n_threads(n);
  float *temp = (float *)malloc(sizeof(float) * (size_t)n);
  float *tbest = (float *)malloc(sizeof(float) * (size_t)bs);
  int *tbesti = (int *)malloc(sizeof(int) * (size_t)bs);
  for (int bi = 0; bi < b; ++bi) {
    const float *pts = dataset + (size_t)bi * n * 3;
    int32_t *out = idxs + (size_t)bi * m;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
      for (int t = 0; t < bs; ++t) { tbest[t] = -1.0f; tbesti[t] = 0; }
      for (int k = 0; k < n; ++k) {
        int t = k % bs;
        float x2 = pts[k * 3 + 0], y2 = pts[k * 3 + 1], z2 = pts[k * 3 + 2];
        float mag = y2 * y2;
        mag = fmaf(x2, x2, mag);
        mag = fmaf(z2, z2, mag);
        if ((double)mag <= 1e-3) continue;
        float d = sqdist3(x2, y2, z2, x1, y1, z1);
        float d2 = fminf(d, temp[k]);
        temp[k] = d2;
        if (d2 > tbest[t]) { tbest[t] = d2; tbesti[t] = k; }
      }
      /* halving tree: on equal values keep the lower tid's entry */
      for (int off = bs / 2; off >= 1; off >>= 1) {
        for (int t = 0; t < off; ++t) {
          float v1 = tbest[t], v2 = tbest[t + off];
          int i1 = tbesti[t], i2 = tbesti[t + off];
          tbest[t] = v1 > v2 ? v1 : v2;
          tbesti[t] = v2 > v1 ? i2 : i1;
        }
      }
      old = tbesti[0];
      out[j] = old;
    }
  }
  free(temp); free(tbest); free(tbesti);
}

/* sampling_gpu.cu:8-20 gather_points_kernel: out[b,c,j] = points[b,c,idx[b,j]]. */
void p2r_ref_gather_points(int b, int c, int n, int m, const float *points, const int32_t *idx,
                           float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}

/* sampling_gpu.cu:34-47 gather_points_grad_kernel: scatter-add into zeros (sampling.cpp:49-51).
 * The reference's atomicAdd order is unspecified; this file adds in index order. */
void p2r_ref_gather_points_grad(int b, int c, int n, int m, const float *grad_out,
                                const int32_t *idx, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j)
        grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] +=
            grad_out[((size_t)i * c + l) * m + j];
}

/* ball_query_gpu.cu:9-44 query_ball_point_kernel; idx zero-initialised (ball_query.cpp:19-21).
 * First nsample points in INDEX order with d2 < radius^2 (strict, fp32 radius2 :22); the first
 * hit pre-fills every slot (:34-38); no hit leaves zeros. */
void p2r_ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                        const float *xyz, int32_t *idx) {
  float radius2 = radius * radius;
  memset(idx, 0, sizeof(int32_t) * (size_t)b * m * nsample);
  for (int bi = 0; bi < b; ++bi) {
    const float *p = xyz + (size_t)bi * n * 3;
    const float *q = new_xyz + (size_t)bi * m * 3;
    int32_t *o = idx + (size_t)bi * m * nsample;
    for (int j = 0; j < m; ++j) {
      float nx = q[j * 3 + 0], ny = q[j * 3 + 1], nz = q[j * 3 + 2];
      int cnt = 0;
      for (int k = 0; k < n && cnt < nsample; ++k) {
        float d2 = sqdist3(nx, ny, nz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[j * nsample + l] = k;
          o[j * nsample + cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* group_points_gpu.cu:8-28: out[b,c,j,k] = points[b,c,idx[b,j,k]]. */
void p2r_ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                          const int32_t *idx, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int ii = idx[((size_t)bi * npoints + j) * nsample + k];
          out[(((size_t)bi * c + l) * npoints + j) * nsample + k] = points[((size_t)bi * c + l) * n + ii];
        }
}

/* group_points_gpu.cu:43-64: scatter-add into zeros (group_points.cpp:48-50). */
void p2r_ref_group_points_grad(int b, int c, int n, int npoints, int nsample,
                               const float *grad_out, const int32_t *idx, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int ii = idx[((size_t)bi * npoints + j) * nsample + k];
          grad_points[((size_t)bi * c + l) * n + ii] +=
              grad_out[(((size_t)bi * c + l) * npoints + j) * nsample + k];
        }
}

/* interpolate_gpu.cu:9-59 three_nn_kernel: running top-3 with double best (1e40) against a float
 * candidate, strict < cascade => equal distances keep the earlier index; writes dist^2 (the
 * Python wrapper takes sqrt, pointnet2_utils.py:124-125). Unused slots: (float)1e40 = +inf, idx 0. */
void p2r_ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                      int32_t *idx) {
  for (int bi = 0; bi < b; ++bi) {
    const float *u = unknown + (size_t)bi * n * 3;
    const float *kn = known + (size_t)bi * m * 3;
    for (int j = 0; j < n; ++j) {
      float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        float d = sqdist3(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      size_t o = ((size_t)bi * n + j) * 3;
      dist2[o + 0] = (float)best1; dist2[o + 1] = (float)best2; dist2[o + 2] = (float)best3;
      idx[o + 0] = besti1; idx[o + 1] = besti2; idx[o + 2] = besti3;
    }
  }
}

/* interpolate_gpu.cu:72-101: out[b,c,j] = p[i1]*w1 + p[i2]*w2 + p[i3]*w3
 * (nvcc: t = p2*w2; t = fma(p1,w1,t); t = fma(p3,w3,t)). */
void p2r_ref_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx,
                               const float *weight, float *out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        size_t o = ((size_t)bi * n + j) * 3;
        const float *p = points + ((size_t)bi * c + l) * m;
        float t = p[idx[o + 1]] * weight[o + 1];
        t = fmaf(p[idx[o + 0]], weight[o + 0], t);
        t = fmaf(p[idx[o + 2]], weight[o + 2], t);
        out[((size_t)bi * c + l) * n + j] = t;
      }
}

/* interpolate_gpu.cu:116-143: three scatter-adds of grad*w into zeros (interpolate.cpp:85-87). */
void p2r_ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                    const int32_t *idx, const float *weight, float *grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        size_t o = ((size_t)bi * n + j) * 3;
        float g = grad_out[((size_t)bi * c + l) * n + j];
        float *gp = grad_points + ((size_t)bi * c + l) * m;
        gp[idx[o + 0]] += g * weight[o + 0];
        gp[idx[o + 1]] += g * weight[o + 1];
        gp[idx[o + 2]] += g * weight[o + 2];
      }
}
