// dataloader_ops.cu -- raw samples resident in HBM -> one training / eval batch, in one launch.
//
// Replaces, for a whole batch, the per-sample numpy work of the reference's dataset class
// (ref: models/p2rnet/dataloader.py:31-84 augment_data, :128-131 frame picking, :137-145 dtype casts,
// :148-160 collate_fn): pick `num_frames` of the raw frames, flip / rotate / translate joints and votes,
// split the vote mask off, and write the three big tensors of the `data` dict in their final layout
//     input_joints (B,T,J,3|4) f32, vote_label (B,T,J,9) f32, vote_label_mask (B,T,J) i64.
// The small per-box labels (<= 10 boxes a sample) stay on the host (pose2room_b200/dataloader.py).
//
// HBM-bound: 52 B read + 56 B written per (frame, joint); at B=32, T=1024, J=25 that is 88 MB a batch.
// Layout: a CTA owns P2R_DL_FRAMES consecutive OUTPUT frames of one batch item.  Raw frame rows
// (J*3 and J*10 floats) are staged into shared memory with coalesced loads, each thread transforms one
// joint out of shared memory (stride-3 / stride-10 word accesses: odd strides, or 2-way at most), the results go
// back to shared memory and leave as coalesced row stores -- the output frames of a CTA are contiguous.
#include "p2r_common.cuh"
#include "p2r_b200.h"
#include "augment_math.h"

#define P2R_DL_FRAMES 8
#define P2R_DL_THREADS 256

__global__ void __launch_bounds__(P2R_DL_THREADS)
make_batch_kernel(const float* __restrict__ joints, const float* __restrict__ votes,
                  const long long* __restrict__ frame_start, const int* __restrict__ sample_ids,
                  const double* __restrict__ params, int num_frames, int J, int out_c,
                  float* __restrict__ input_joints, float* __restrict__ vote_label,
                  long long* __restrict__ vote_label_mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * P2R_DL_FRAMES;
  const int nf = min(P2R_DL_FRAMES, num_frames - t0);
  const int rowj = J * 3, rowv = J * 10, rowo = J * out_c, rowl = J * 9;

  long long* s_mask = reinterpret_cast<long long*>(smem_raw);                 // [FR * J]
  float* s_in_j = reinterpret_cast<float*>(s_mask + P2R_DL_FRAMES * J);        // [FR * J*3]
  float* s_in_v = s_in_j + P2R_DL_FRAMES * rowj;                                // [FR * J*10]
  float* s_out_j = s_in_v + P2R_DL_FRAMES * rowv;                               // [FR * J*out_c]
  float* s_out_v = s_out_j + P2R_DL_FRAMES * rowo;                              // [FR * J*9]
  __shared__ double s_p[P2R_AUG_STRIDE];
  __shared__ long long s_src[P2R_DL_FRAMES];

  const int sample = sample_ids[b];
  const long long f0 = frame_start[sample];
  const int n_raw = (int)(frame_start[sample + 1] - f0);
  if (threadIdx.x < P2R_AUG_STRIDE) s_p[threadIdx.x] = params[(size_t)b * P2R_AUG_STRIDE + threadIdx.x];
  if (threadIdx.x < nf) s_src[threadIdx.x] = f0 + p2r_frame_id(n_raw, num_frames, t0 + threadIdx.x);
  __syncthreads();

  // stage the raw rows (each row is contiguous in global memory)
  for (int i = threadIdx.x; i < nf * rowj; i += P2R_DL_THREADS) {
    const int f = i / rowj, c = i - f * rowj;
    s_in_j[i] = __ldg(joints + s_src[f] * rowj + c);
  }
  for (int i = threadIdx.x; i < nf * rowv; i += P2R_DL_THREADS) {
    const int f = i / rowv, c = i - f * rowv;
    s_in_v[i] = __ldg(votes + s_src[f] * rowv + c);
  }
  __syncthreads();

  for (int i = threadIdx.x; i < nf * J; i += P2R_DL_THREADS) {
    float oj[4], ov[9];
    s_mask[i] = p2r_augment_joint(s_in_j + (size_t)i * 3, s_in_v + (size_t)i * 10, s_p, out_c, oj, ov);
    for (int c = 0; c < out_c; ++c) s_out_j[(size_t)i * out_c + c] = oj[c];
#pragma unroll
    for (int c = 0; c < 9; ++c) s_out_v[(size_t)i * 9 + c] = ov[c];
  }
  __syncthreads();

  // the CTA's output frames are contiguous: plain coalesced copies
  const size_t frame0 = (size_t)b * num_frames + t0;
  float* gj = input_joints + frame0 * rowo;
  for (int i = threadIdx.x; i < nf * rowo; i += P2R_DL_THREADS) gj[i] = s_out_j[i];
  float* gv = vote_label + frame0 * rowl;
  for (int i = threadIdx.x; i < nf * rowl; i += P2R_DL_THREADS) gv[i] = s_out_v[i];
  long long* gm = vote_label_mask + frame0 * J;
  for (int i = threadIdx.x; i < nf * J; i += P2R_DL_THREADS) gm[i] = s_mask[i];
}

extern "C" int p2r_make_batch(const float* joints, const float* votes, const long long* frame_start,
                              const int* sample_ids, const double* params, int b, int num_frames, int j,
                              int out_channels, float* input_joints, float* vote_label,
                              long long* vote_label_mask, void* stream) {
  P2R_CHECK_ARG(b >= 0 && num_frames >= 0 && j > 0, "p2r_make_batch");
  P2R_CHECK_ARG(out_channels == 3 || out_channels == 4, "p2r_make_batch");
  P2R_CHECK_ARG(b <= 65535, "p2r_make_batch");
  if (b == 0 || num_frames == 0) return 0;
  const size_t smem = (size_t)P2R_DL_FRAMES * j * (sizeof(long long) + sizeof(float) * (3 + 10 + out_channels + 9));
  P2R_CHECK_ARG(smem <= 200 * 1024, "p2r_make_batch");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(make_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { p2r_set_last_error("p2r_make_batch", (int)e); return (int)e; }
  }
  dim3 grid(p2r_ceil_div(num_frames, P2R_DL_FRAMES), b);
  make_batch_kernel<<<grid, P2R_DL_THREADS, smem, (cudaStream_t)stream>>>(
      joints, votes, frame_start, sample_ids, params, num_frames, j, out_channels, input_joints, vote_label,
      vote_label_mask);
  P2R_RETURN_LAUNCH("p2r_make_batch");
}
