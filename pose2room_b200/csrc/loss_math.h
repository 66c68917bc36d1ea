/*
 * loss_math.h -- per-seed / per-proposal arithmetic of the fused detection-loss kernel (loss_ops.cu).
 *
 * Restates BoxNetDetectionLoss (ref: models/loss.py:42-189) element by element:
 *   compute_vote_loss            :90-115   p2rl_seed()
 *   compute_correspondence       :117-150  p2rl_proposal(), first half
 *   compute_box_and_sem_cls_loss :42-88    p2rl_proposal(), second half + the per-ground-truth side (p2rl_gt_side())
 *   __call__                     :152-189  p2rl_finalize()
 * Everything that DECIDES something (nearest ground truth, near / far thresholds, the vote target, argmax of the
 * objectness scores) is evaluated in float32 with the rounding order of the reference's torch expressions
 * (net_utils/nn_distance.py:49-52: diff, diff**2 and the sum over 3 are separate float32 roundings; ties go to the
 * first index like torch.min / argmin), so labels and assignments are the reference's.  The sums behind the ten
 * reported numbers are accumulated in float64 (the reference: float32 torch.sum; agreement ~1e-7 relative).
 *
 * Gradients: every differentiable input feeds exactly one loss term, and every term is  sum / (count + 1e-6).  The
 * forward therefore stores the UN-NORMALISED gradient of each sum (u_*) plus the four reciprocals (scales); the
 * backward is one elementwise pass  grad = upstream(term) * scale * u.
 *
 * The same source compiles for the device (nvcc) and for the host (gcc -ffp-contract=off; used only by
 * tests/test_loss_math.py to hold this arithmetic to the CPU oracle without a GPU).
 */
#ifndef P2R_LOSS_MATH_H
#define P2R_LOSS_MATH_H

#ifdef __CUDACC__
#define P2RL_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define P2RL_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define P2RL_FADD(a, b) __fadd_rn((a), (b))
#define P2RL_FSUB(a, b) __fsub_rn((a), (b))
#define P2RL_FMUL(a, b) __fmul_rn((a), (b))
#define P2RL_FSQRT(a) __fsqrt_rn(a)
#else
#define P2RL_FADD(a, b) ((float)((float)(a) + (float)(b)))
#define P2RL_FSUB(a, b) ((float)((float)(a) - (float)(b)))
#define P2RL_FMUL(a, b) ((float)((float)(a) * (float)(b)))
#define P2RL_FSQRT(a) sqrtf(a)
#endif

#define P2RL_GT_VOTE_FACTOR 3     /* loss.py:13 */
#define P2RL_NEAR 0.3f            /* loss.py:11, compared in float32 like `tensor < 0.3` */
#define P2RL_FAR 0.6f             /* loss.py:10 */
#define P2RL_OBJ_W0 0.1f          /* loss.py:14 OBJECTNESS_CLS_WEIGHTS */
#define P2RL_OBJ_W1 0.9f
#define P2RL_MAX_GT 64            /* ground-truth slots per sample the kernel stages (the dataset has 10) */

/* partial sums every block / the host loop accumulates (float64) */
enum {
  P2RL_S_VOTE = 0,   /* sum_seeds  mean_3 huber(vote - target) * mask                 */
  P2RL_S_VOTE_CNT,   /* sum_seeds  mask                                                */
  P2RL_S_OBJ,        /* sum_props  weighted CE(objectness) * objectness_mask           */
  P2RL_S_OBJMASK,    /* sum_props  objectness_mask                                     */
  P2RL_S_POS,        /* sum_props  objectness_label                                    */
  P2RL_S_C1,         /* sum_props  min_g |center - gt_g|^2 * label                     */
  P2RL_S_SIZE,       /* sum_props  mean_3 huber(size - gt_size[assign]) * label        */
  P2RL_S_HEAD,       /* sum_props  mean_2 huber(heading - gt_heading[assign]) * label  */
  P2RL_S_SEM,        /* sum_props  CE(sem_cls, gt_cls[assign]) * label                 */
  P2RL_S_C2,         /* sum_gt     min_p |center_p - gt|^2 * box_label_mask            */
  P2RL_S_BOXMASK,    /* sum_gt     box_label_mask                                      */
  P2RL_S_ACC,        /* sum_props  (argmax(objectness) == label) * objectness_mask     */
  P2RL_NSUM
};

/* float32 outputs, float64 outputs, scales */
enum { P2RL_O_VOTE = 0, P2RL_O_OBJ, P2RL_O_CENTER, P2RL_O_SIZE, P2RL_O_SEM, P2RL_O_POS_RATIO, P2RL_O_NEG_RATIO,
       P2RL_O_OBJ_ACC, P2RL_NOUT32 };
enum { P2RL_O_HEADING = 0, P2RL_O_TOTAL, P2RL_NOUT64 };
enum { P2RL_SC_VOTE = 0, P2RL_SC_OBJMASK, P2RL_SC_POS, P2RL_SC_BOXMASK, P2RL_NSCALE };

/* |a - b|^2 as torch evaluates sum((a - b) ** 2, -1) over 3 float32 components */
P2RL_HD float p2rl_sqdist3(const float* a, const float* b) {
  const float d0 = P2RL_FSUB(a[0], b[0]), d1 = P2RL_FSUB(a[1], b[1]), d2 = P2RL_FSUB(a[2], b[2]);
  return P2RL_FADD(P2RL_FADD(P2RL_FMUL(d0, d0), P2RL_FMUL(d1, d1)), P2RL_FMUL(d2, d2));
}

/* huber_loss(e, delta=1) (net_utils/nn_distance.py:15-32) and its derivative, in the type of e */
P2RL_HD float p2rl_huber_f(float e, float* de) {
  const float a = fabsf(e), q = a < 1.0f ? a : 1.0f;
  *de = a <= 1.0f ? e : (e > 0.0f ? 1.0f : -1.0f);
  return P2RL_FADD(P2RL_FMUL(0.5f, P2RL_FMUL(q, q)), P2RL_FSUB(a, q));
}
P2RL_HD double p2rl_huber_d(double e, double* de) {
  const double a = fabs(e), q = a < 1.0 ? a : 1.0;
  *de = a <= 1.0 ? e : (e > 0.0 ? 1.0 : -1.0);
  return 0.5 * q * q + (a - q);
}

/* One seed (loss.py:90-115).  sk: the seed's skeleton, J x 3; gv: vote_label[b, seed_ind, origin, 0:9]; mask:
 * vote_label_mask[b, seed_ind, origin]; v: the predicted vote.  u_vote[3]: un-normalised d(sum)/d(v). */
P2RL_HD void p2rl_seed(const float* sk, int J, int origin, const float* gv, long long mask, const float* v,
                       float* u_vote, double* sums) {
  float votes[P2RL_GT_VOTE_FACTOR][3];
  for (int s = 0; s < P2RL_GT_VOTE_FACTOR; ++s)
    for (int c = 0; c < 3; ++c) votes[s][c] = P2RL_FADD(sk[origin * 3 + c], gv[s * 3 + c]);
  /* dist2[j] = min_s |votes_s - sk_j|^2 (first s), then the first minimal j; pick = ind2[j] */
  float best = 0.0f;
  int pick = 0;
  for (int j = 0; j < J; ++j) {
    float dj = 0.0f;
    int sj = 0;
    for (int s = 0; s < P2RL_GT_VOTE_FACTOR; ++s) {
      const float d = p2rl_sqdist3(votes[s], sk + j * 3);
      if (s == 0 || d < dj) { dj = d; sj = s; }
    }
    if (j == 0 || dj < best) { best = dj; pick = sj; }
  }
  const float m = (float)mask;
  double h = 0.0;
  for (int c = 0; c < 3; ++c) {
    float de;
    h += (double)p2rl_huber_f(P2RL_FSUB(v[c], votes[pick][c]), &de);
    u_vote[c] = m * de * (1.0f / 3.0f);
  }
  sums[P2RL_S_VOTE] += h / 3.0 * (double)m;
  sums[P2RL_S_VOTE_CNT] += (double)m;
}

/* One proposal (loss.py:117-150 and :42-88 without the per-ground-truth side of the centre loss).
 * Ground truth of the proposal's sample: gt_center [G,3], gt_mask [G], gt_size [G,3], gt_heading [G,2], gt_cls [G].
 * heading / u_head are float64 when heading_f64 (the reference's heading mixture is float64, mdn.py:28) else float32.
 * Writes the un-normalised gradients u_c1[3], u_size[3], u_head[2], u_obj[2], u_sem[C]. */
P2RL_HD void p2rl_proposal(const float* agg, const float* center, const float* size, const void* heading,
                           int heading_f64, const float* obj, const float* sem, int C, const float* gt_center,
                           const float* gt_mask, const float* gt_size, const float* gt_heading,
                           const long long* gt_cls, int G, float* u_c1, float* u_size, void* u_head, float* u_obj,
                           float* u_sem, double* sums) {
  /* nearest VALID ground-truth centre; the reference indexes the compacted list per_gt_center[per_mask > 0] and
   * then gathers from the un-compacted arrays with that index (loss.py:128,68) -- same here */
  float d1 = 0.0f;
  int assign = 0, rank = 0;
  for (int g = 0; g < G; ++g) {
    if (!(gt_mask[g] > 0.0f)) continue;
    const float d = p2rl_sqdist3(agg, gt_center + g * 3);
    if (rank == 0 || d < d1) { d1 = d; assign = rank; }
    ++rank;
  }
  /* a sample without a valid box (the reference would raise on the empty min): far from everything */
  const float eu = rank ? P2RL_FSQRT(P2RL_FADD(d1, 1e-6f)) : 3.0e38f;
  const int label = eu < P2RL_NEAR;
  const float omask = (label || eu > P2RL_FAR) ? 1.0f : 0.0f;
  const float lab = (float)label;

  /* objectness: weighted cross entropy over 2 logits */
  {
    const float mx = obj[0] > obj[1] ? obj[0] : obj[1];
    const float e0 = expf(obj[0] - mx), e1 = expf(obj[1] - mx);
    const float se = e0 + e1, lse = mx + logf(se);
    const float w = label ? P2RL_OBJ_W1 : P2RL_OBJ_W0;
    const float ce = w * (lse - obj[label]);
    u_obj[0] = omask * w * (e0 / se - (label == 0 ? 1.0f : 0.0f));
    u_obj[1] = omask * w * (e1 / se - (label == 1 ? 1.0f : 0.0f));
    const int pred = obj[1] > obj[0];          /* torch.argmax: first maximum */
    sums[P2RL_S_OBJ] += (double)ce * (double)omask;
    sums[P2RL_S_OBJMASK] += (double)omask;
    sums[P2RL_S_POS] += (double)lab;
    sums[P2RL_S_ACC] += (pred == label) ? (double)omask : 0.0;
  }
  /* centre, proposal side: nearest of ALL G slots (padded slots included, like loss.py:64) */
  {
    float dc = 0.0f;
    int ic = 0;
    for (int g = 0; g < G; ++g) {
      const float d = p2rl_sqdist3(center, gt_center + g * 3);
      if (g == 0 || d < dc) { dc = d; ic = g; }
    }
    sums[P2RL_S_C1] += (double)dc * (double)lab;
    for (int c = 0; c < 3; ++c) u_c1[c] = lab * 2.0f * P2RL_FSUB(center[c], gt_center[ic * 3 + c]);
  }
  /* size */
  {
    double h = 0.0;
    for (int c = 0; c < 3; ++c) {
      float de;
      h += (double)p2rl_huber_f(P2RL_FSUB(size[c], gt_size[assign * 3 + c]), &de);
      u_size[c] = lab * de * (1.0f / 3.0f);
    }
    sums[P2RL_S_SIZE] += h / 3.0 * (double)lab;
  }
  /* heading (sin, cos) */
  {
    double h = 0.0;
    for (int c = 0; c < 2; ++c) {
      const double gth = (double)gt_heading[assign * 2 + c];
      if (heading_f64) {
        double de;
        h += p2rl_huber_d(((const double*)heading)[c] - gth, &de);
        ((double*)u_head)[c] = (double)lab * de * 0.5;
      } else {
        float de;
        h += (double)p2rl_huber_f(P2RL_FSUB(((const float*)heading)[c], gt_heading[assign * 2 + c]), &de);
        ((float*)u_head)[c] = lab * de * 0.5f;
      }
    }
    sums[P2RL_S_HEAD] += h * 0.5 * (double)lab;
  }
  /* semantic class: cross entropy over C logits against gt_cls[assign] */
  {
    const long long cls = gt_cls[assign];
    float mx = sem[0];
    for (int k = 1; k < C; ++k) mx = sem[k] > mx ? sem[k] : mx;
    float se = 0.0f;
    for (int k = 0; k < C; ++k) se += expf(sem[k] - mx);
    const float lse = mx + logf(se);
    const int ok = cls >= 0 && cls < C;       /* torch raises on an out-of-range target; here the term is dropped */
    if (ok) sums[P2RL_S_SEM] += (double)(lse - sem[cls]) * (double)lab;
    for (int k = 0; k < C; ++k)
      u_sem[k] = ok ? lab * (expf(sem[k] - mx) / se - (k == (int)cls ? 1.0f : 0.0f)) : 0.0f;
  }
}

/* Ground-truth side of the centre loss for ONE ground-truth slot g whose nearest proposal (first minimum over p of
 * |center_p - gt_g|^2) is already known: adds to the sums; the caller adds  mask * 2 * (center_p - gt_g)  to u_c2[p]. */
P2RL_HD void p2rl_gt_side(float d2, float mask, double* sums) {
  sums[P2RL_S_C2] += (double)d2 * (double)mask;
  sums[P2RL_S_BOXMASK] += (double)mask;
}

/* loss.py:152-189: the ten reported numbers + the reciprocals the backward needs. */
P2RL_HD void p2rl_finalize(const double* s, double n_proposals, float* out32, double* out64, double* scales) {
  const double sv = 1.0 / (s[P2RL_S_VOTE_CNT] + 1e-6), so = 1.0 / (s[P2RL_S_OBJMASK] + 1e-6);
  const double sp = 1.0 / (s[P2RL_S_POS] + 1e-6), sb = 1.0 / (s[P2RL_S_BOXMASK] + 1e-6);
  const double vote = s[P2RL_S_VOTE] * sv, objn = s[P2RL_S_OBJ] * so;
  const double center = 0.5 * (s[P2RL_S_C1] * sp + s[P2RL_S_C2] * sb);
  const double size = s[P2RL_S_SIZE] * sp, head = s[P2RL_S_HEAD] * sp, sem = s[P2RL_S_SEM] * sp;
  const double pos = s[P2RL_S_POS] / n_proposals;
  out32[P2RL_O_VOTE] = (float)vote;
  out32[P2RL_O_OBJ] = (float)objn;
  out32[P2RL_O_CENTER] = (float)center;
  out32[P2RL_O_SIZE] = (float)size;
  out32[P2RL_O_SEM] = (float)sem;
  out32[P2RL_O_POS_RATIO] = (float)pos;
  out32[P2RL_O_NEG_RATIO] = (float)(s[P2RL_S_OBJMASK] / n_proposals - pos);
  out32[P2RL_O_OBJ_ACC] = (float)(s[P2RL_S_ACC] * so);
  out64[P2RL_O_HEADING] = head;
  out64[P2RL_O_TOTAL] = 10.0 * vote + 5.0 * objn + 10.0 * center + 10.0 * size + 10.0 * head + sem;
  scales[P2RL_SC_VOTE] = sv;
  scales[P2RL_SC_OBJMASK] = so;
  scales[P2RL_SC_POS] = sp;
  scales[P2RL_SC_BOXMASK] = sb;
}

/* Upstream gradients of the reported numbers -> effective weight of each of the six differentiable terms
 * (total = 10 vote + 5 objectness + 10 center + 10 size + 10 heading + sem, loss.py:168). */
enum { P2RL_T_VOTE = 0, P2RL_T_OBJ, P2RL_T_CENTER, P2RL_T_SIZE, P2RL_T_HEADING, P2RL_T_SEM, P2RL_NTERM };
P2RL_HD void p2rl_term_weights(const float* g32, const double* g64, double* w) {
  const double gt = g64[P2RL_O_TOTAL];
  w[P2RL_T_VOTE] = (double)g32[P2RL_O_VOTE] + 10.0 * gt;
  w[P2RL_T_OBJ] = (double)g32[P2RL_O_OBJ] + 5.0 * gt;
  w[P2RL_T_CENTER] = (double)g32[P2RL_O_CENTER] + 10.0 * gt;
  w[P2RL_T_SIZE] = (double)g32[P2RL_O_SIZE] + 10.0 * gt;
  w[P2RL_T_HEADING] = g64[P2RL_O_HEADING] + 10.0 * gt;
  w[P2RL_T_SEM] = (double)g32[P2RL_O_SEM] + gt;
}

#endif /* P2R_LOSS_MATH_H */
