/* TEST HARNESS (CPU): runs the product's detection-loss arithmetic header (pose2room_b200/csrc/loss_math.h) over a
 * batch with the same indexing, the same first-minimum rules and the same sum -> finalize -> scale structure as
 * detection_loss_kernel / detection_loss_grad_kernel (csrc/loss_ops.cu), so tests/test_loss_math.py can hold that
 * arithmetic to the CPU oracle (oracle/model_ref.py, pinned by the reference's goldens) without a GPU.  Built by the
 * test with gcc -ffp-contract=off; never shipped, never loaded by pose2room_b200/. */
#include <stddef.h>
#include <string.h>
#include "loss_math.h"

void host_detection_loss(const float* vote_xyz, const float* center, const float* size, const void* heading,
                         int heading_f64, const float* obj, int obj_stride, const float* sem, int sem_stride,
                         const float* agg, const float* skeleton, const long long* seed_inds, const float* vote_label,
                         const long long* vote_mask, const float* gt_center, const float* gt_mask, const float* gt_size,
                         const float* gt_heading, const long long* gt_cls, int B, int S, int J, int T, int P, int G, int C,
                         int origin, float* out32, double* out64, double* scales, float* u_vote, float* u_c1,
                         float* u_c2, float* u_size, void* u_head, float* u_obj, float* u_sem) {
  double sums[P2RL_NSUM];
  for (int k = 0; k < P2RL_NSUM; ++k) sums[k] = 0.0;
  for (int b = 0; b < B; ++b) {
    const float* gc = gt_center + (size_t)b * G * 3;
    const float* gm = gt_mask + (size_t)b * G;
    for (int p = 0; p < P; ++p) {
      const size_t r = (size_t)b * P + p;
      const void* hd = heading_f64 ? (const void*)((const double*)heading + r * 2) : (const void*)((const float*)heading + r * 2);
      void* uh = heading_f64 ? (void*)((double*)u_head + r * 2) : (void*)((float*)u_head + r * 2);
      p2rl_proposal(agg + r * 3, center + r * 3, size + r * 3, hd, heading_f64, obj + r * obj_stride, sem + r * sem_stride,
                    C, gc, gm, gt_size + (size_t)b * G * 3, gt_heading + (size_t)b * G * 2, gt_cls + (size_t)b * G, G,
                    u_c1 + r * 3, u_size + r * 3, uh, u_obj + r * 2, u_sem + r * C, sums);
      for (int c = 0; c < 3; ++c) u_c2[r * 3 + c] = 0.0f;
    }
    for (int g = 0; g < G; ++g) {
      float best = 0.0f;
      int bi = -1;
      for (int p = 0; p < P; ++p) {
        const float d = p2rl_sqdist3(center + ((size_t)b * P + p) * 3, gc + g * 3);
        if (bi < 0 || d < best) { best = d; bi = p; }
      }
      p2rl_gt_side(best, gm[g], sums);
      const size_t r = (size_t)b * P + bi;
      for (int c = 0; c < 3; ++c)
        u_c2[r * 3 + c] = P2RL_FADD(u_c2[r * 3 + c], gm[g] * 2.0f * P2RL_FSUB(center[r * 3 + c], gc[g * 3 + c]));
    }
  }
  for (long long seed = 0; seed < (long long)B * S; ++seed) {
    const int b = (int)(seed / S);
    const size_t row = ((size_t)b * T + (size_t)seed_inds[seed]) * J + origin;
    p2rl_seed(skeleton + (size_t)seed * J * 3, J, origin, vote_label + row * 9, vote_mask[row], vote_xyz + (size_t)seed * 3,
              u_vote + (size_t)seed * 3, sums);
  }
  p2rl_finalize(sums, (double)B * (double)P, out32, out64, scales);
}

void host_detection_loss_grad(const float* g32, const double* g64, const double* scales, const float* u_vote,
                              const float* u_c1, const float* u_c2, const float* u_size, const void* u_head,
                              int heading_f64, const float* u_obj, const float* u_sem, int B, int S, int P, int C,
                              float* d_vote, float* d_center, float* d_size, void* d_head, float* d_obj, float* d_sem) {
  double w[P2RL_NTERM];
  p2rl_term_weights(g32, g64, w);
  const double sv = scales[P2RL_SC_VOTE], so = scales[P2RL_SC_OBJMASK], sp = scales[P2RL_SC_POS], sb = scales[P2RL_SC_BOXMASK];
  for (long long i = 0; i < (long long)B * S * 3; ++i) d_vote[i] = (float)(w[P2RL_T_VOTE] * sv * (double)u_vote[i]);
  for (long long i = 0; i < (long long)B * P * 3; ++i) {
    d_center[i] = (float)(w[P2RL_T_CENTER] * 0.5 * (sp * (double)u_c1[i] + sb * (double)u_c2[i]));
    d_size[i] = (float)(w[P2RL_T_SIZE] * sp * (double)u_size[i]);
  }
  for (long long i = 0; i < (long long)B * P * 2; ++i) {
    if (heading_f64) ((double*)d_head)[i] = w[P2RL_T_HEADING] * sp * ((const double*)u_head)[i];
    else ((float*)d_head)[i] = (float)(w[P2RL_T_HEADING] * sp * (double)((const float*)u_head)[i]);
    d_obj[i] = (float)(w[P2RL_T_OBJ] * so * (double)u_obj[i]);
  }
  for (long long i = 0; i < (long long)B * P * C; ++i) d_sem[i] = (float)(w[P2RL_T_SEM] * sp * (double)u_sem[i]);
}
