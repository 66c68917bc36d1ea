/* TEST HARNESS (CPU): runs the product's per-joint arithmetic header (pose2room_b200/csrc/augment_math.h) over a
 * batch with the same indexing as make_batch_kernel, so tests/test_dataloader_math.py can check that arithmetic
 * against the reference goldens without a GPU.  Built by the test with gcc -ffp-contract=off; never shipped, never
 * loaded by pose2room_b200/. */
#include <stddef.h>
#include "augment_math.h"

void host_make_batch(const float* joints, const float* votes, const long long* frame_start, const int* sample_ids,
                     const double* params, int b, int num_frames, int j, int out_c, float* input_joints,
                     float* vote_label, long long* vote_label_mask) {
  for (int bi = 0; bi < b; ++bi) {
    const long long f0 = frame_start[sample_ids[bi]];
    const int n_raw = (int)(frame_start[sample_ids[bi] + 1] - f0);
    for (int t = 0; t < num_frames; ++t) {
      const long long src = f0 + p2r_frame_id(n_raw, num_frames, t);
      for (int k = 0; k < j; ++k) {
        const size_t o = ((size_t)bi * num_frames + t) * j + k;
        vote_label_mask[o] = p2r_augment_joint(joints + (src * j + k) * 3, votes + (src * j + k) * 10,
                                               params + (size_t)bi * P2R_AUG_STRIDE, out_c, input_joints + o * out_c,
                                               vote_label + o * 9);
      }
    }
  }
}

int host_frame_id(int n_raw, int num_frames, int t) { return p2r_frame_id(n_raw, num_frames, t); }
