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
#include <stdlib.h>

#define P2R_DL_FRAMES 8
#define P2R_DL_THREADS 256

__global__ void __launch_bounds__(P2R_DL_THREADS)
make_batch_kernel(const float* __restrict__ joints, const float* __restrict__ votes,
                  const long long* __restrict__ frame_start, const int* __restrict__ sample_ids,
                  const double* __restrict__ params, int num_frames, int J, int out_c,
                  float* __restrict__ input_joints, float* __restrict__ vote_label,
                  long long* __restrict__ vote_label_mask) {
  P2R_DYN_SMEM(unsigned char, smem_raw);
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

// ---------------------------------------------------------------------------------------------------------------
// Variant 2 (the default since round 2; P2R_MAKE_BATCH_VARIANT=1 selects variant 1; bench.py's data_path.variants A/B).
// Same arithmetic, different data movement: persistent CTAs walk over (batch item, group of 8 frames) work items with
// a 3-stage cp.async ring, so the raw rows of the next two groups are in flight (no registers, no thread waiting on
// them) while the current group is transformed and stored -- variant 1 exposes the full load latency of every CTA.
// One warp per frame: a warp copies its frame's two rows (joints in 4-byte, votes in 8-byte pieces: raw frames are
// only 4- / 8-byte aligned), then transforms that frame's joints.  Only __syncthreads and cp.async.wait_group: there
// is no spin-wait in this kernel.
#define P2R_DL_STAGES 3

#ifdef P2R_HOST_EMULATION
// host emulator: the copy happens at issue time (the strictest schedule: a stage that is overwritten while another
// thread still reads it, or read before the matching wait + barrier, shows up as a data race / a wrong value)
__device__ __forceinline__ void p2r_cp_async4(void* dst, const void* src) { memcpy(dst, src, 4); }
__device__ __forceinline__ void p2r_cp_async8(void* dst, const void* src) { memcpy(dst, src, 8); }
__device__ __forceinline__ void p2r_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
__device__ __forceinline__ void p2r_cp_async_commit() {}
template <int N>
__device__ __forceinline__ void p2r_cp_async_wait() {}
#else
__device__ __forceinline__ void p2r_cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(p2r_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void p2r_cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(p2r_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void p2r_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(p2r_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void p2r_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void p2r_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// One warp copies one raw row (`bytes`, a multiple of 4; source 4-byte aligned) to `dst`, which the caller places at the
// SAME offset modulo 16 as the source: up to three leading words, then whole 16-byte pieces, then up to three trailing words
// -- a quarter of the copy instructions of a word-by-word (or half of an 8-byte) copy.
__device__ __forceinline__ void p2r_warp_row_copy(unsigned char* dst, const unsigned char* src, int bytes, int lane) {
  const int head = (int)((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15);     // bytes up to the first 16-byte boundary
  const int hb = head < bytes ? head : bytes;
  const int body = (bytes - hb) >> 4;
  const int tail0 = hb + (body << 4);
  if (lane < (hb >> 2)) p2r_cp_async4(dst + 4 * lane, src + 4 * lane);
  for (int c = lane; c < body; c += 32) p2r_cp_async16(dst + hb + 16 * c, src + hb + 16 * c);
  if (lane < ((bytes - tail0) >> 2)) p2r_cp_async4(dst + tail0 + 4 * lane, src + tail0 + 4 * lane);
}

__global__ void __launch_bounds__(P2R_DL_THREADS)
make_batch_pipe_kernel(const float* __restrict__ joints, const float* __restrict__ votes,
                       const long long* __restrict__ frame_start, const int* __restrict__ sample_ids,
                       const double* __restrict__ params, int num_frames, int J, int out_c, int groups_per_item,
                       int total_groups, float* __restrict__ input_joints, float* __restrict__ vote_label,
                       long long* __restrict__ vote_label_mask) {
  static_assert(P2R_DL_THREADS == 32 * P2R_DL_FRAMES, "one warp per frame of a group");
  P2R_DYN_SMEM(unsigned char, smem_raw);
  const int rowj = J * 3, rowv = J * 10, rowo = J * out_c, rowl = J * 9;
  // layout: mask | stage[3] { votes slots | joint slots } | out joints | out votes.  A slot is a raw row + 16 bytes of
  // slack, rounded up to 16 bytes: the row is placed at (source address mod 16) inside it, so source and destination
  // are congruent and the copy can use 16-byte pieces (p2r_warp_row_copy)
  long long* s_mask = reinterpret_cast<long long*>(smem_raw);                      // [FR * J], padded to 16 bytes
  const int mask_floats = ((P2R_DL_FRAMES * J * 2 + 3) / 4) * 4;
  float* s_stage = reinterpret_cast<float*>(smem_raw) + mask_floats;                // 16-byte aligned
  const int slotv = ((rowv + 4 + 3) / 4) * 4, slotj = ((rowj + 4 + 3) / 4) * 4;     // floats per slot
  const int stage_floats = P2R_DL_FRAMES * (slotv + slotj);
  float* s_out_j = s_stage + P2R_DL_STAGES * stage_floats;
  float* s_out_v = s_out_j + P2R_DL_FRAMES * rowo;
  __shared__ double s_p[P2R_DL_STAGES][P2R_AUG_STRIDE];
  __shared__ long long s_src[P2R_DL_STAGES][P2R_DL_FRAMES];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the shared-memory sources of the vector stores must be 16-byte aligned too (s_mask is: it is the base; the stage ring
  // is 8-byte granular, so s_out_j / s_out_v may sit on an odd 8-byte boundary for some J)
  const bool out_vec16 = ((reinterpret_cast<uintptr_t>(s_out_j) | reinterpret_cast<uintptr_t>(s_out_v) |
                           reinterpret_cast<uintptr_t>(s_mask)) & 15) == 0;

  // (1) per work item: source frames + parameter block into slot `slot` (threads 0..7 and 32..47)
  auto prepare = [&](int w, int slot) {
    if (w >= total_groups) return;
    const int b = w / groups_per_item, t0 = (w - b * groups_per_item) * P2R_DL_FRAMES;
    if (threadIdx.x < P2R_DL_FRAMES && t0 + (int)threadIdx.x < num_frames) {
      const int sample = sample_ids[b];
      const long long f0 = frame_start[sample];
      const int n_raw = (int)(frame_start[sample + 1] - f0);
      s_src[slot][threadIdx.x] = f0 + p2r_frame_id(n_raw, num_frames, t0 + threadIdx.x);
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + P2R_AUG_STRIDE)
      s_p[slot][threadIdx.x - 32] = params[(size_t)b * P2R_AUG_STRIDE + (threadIdx.x - 32)];
  };
  // (2) asynchronous copy of the raw rows of item w into stage `slot`: warp f copies frame f
  auto issue = [&](int w, int slot) {
    if (w < total_groups) {
      const int b = w / groups_per_item, t0 = (w - b * groups_per_item) * P2R_DL_FRAMES;
      if (t0 + warp < num_frames) {
        const long long src = s_src[slot][warp];
        const unsigned char* gv = reinterpret_cast<const unsigned char*>(votes + src * rowv);    // 40 J bytes per frame
        const unsigned char* gj = reinterpret_cast<const unsigned char*>(joints + src * rowj);   // 12 J bytes per frame
        unsigned char* dv = reinterpret_cast<unsigned char*>(s_stage + slot * stage_floats + warp * slotv) +
                            (reinterpret_cast<uintptr_t>(gv) & 15);
        unsigned char* dj = reinterpret_cast<unsigned char*>(s_stage + slot * stage_floats + P2R_DL_FRAMES * slotv + warp * slotj) +
                            (reinterpret_cast<uintptr_t>(gj) & 15);
        p2r_warp_row_copy(dv, gv, rowv * 4, lane);
        p2r_warp_row_copy(dj, gj, rowj * 4, lane);
      }
    }
    p2r_cp_async_commit();     // always: the group count per iteration stays uniform
  };

  const int w0 = blockIdx.x, wstep = gridDim.x;
#pragma unroll
  for (int s = 0; s < P2R_DL_STAGES - 1; ++s) prepare(w0 + s * wstep, s);
  __syncthreads();
#pragma unroll
  for (int s = 0; s < P2R_DL_STAGES - 1; ++s) issue(w0 + s * wstep, s);

  int k = 0;
  for (int w = w0; w < total_groups; w += wstep, ++k) {
    const int cur = k % P2R_DL_STAGES, nxt = (k + P2R_DL_STAGES - 1) % P2R_DL_STAGES;
    prepare(w + (P2R_DL_STAGES - 1) * wstep, nxt);
    __syncthreads();                               // slot nxt: written above, last read two syncs ago (item k-1)
    issue(w + (P2R_DL_STAGES - 1) * wstep, nxt);
    p2r_cp_async_wait<P2R_DL_STAGES - 1>();        // this thread's copies of item k have landed ...
    __syncthreads();                               // ... and so have everybody else's
    const int b = w / groups_per_item, t0 = (w - b * groups_per_item) * P2R_DL_FRAMES;
    const int nf = min(P2R_DL_FRAMES, num_frames - t0);
    if (warp < nf) {
      const long long src = s_src[cur][warp];      // (the row sits at its source's offset modulo 16 inside its slot)
      const float* in_v = s_stage + cur * stage_floats + warp * slotv +
                          ((reinterpret_cast<uintptr_t>(votes + src * rowv) & 15) >> 2);
      const float* in_j = s_stage + cur * stage_floats + P2R_DL_FRAMES * slotv + warp * slotj +
                          ((reinterpret_cast<uintptr_t>(joints + src * rowj) & 15) >> 2);
      for (int j = lane; j < J; j += 32) {
        float oj[4], ov[9];
        const int i = warp * J + j;
        s_mask[i] = p2r_augment_joint(in_j + j * 3, in_v + j * 10, s_p[cur], out_c, oj, ov);
        for (int c = 0; c < out_c; ++c) s_out_j[(size_t)i * out_c + c] = oj[c];
#pragma unroll
        for (int c = 0; c < 9; ++c) s_out_v[(size_t)i * 9 + c] = ov[c];
      }
    }
    __syncthreads();
    const size_t frame0 = (size_t)b * num_frames + t0;
    float* gj = input_joints + frame0 * rowo;
    float* gv = vote_label + frame0 * rowl;
    long long* gm = vote_label_mask + frame0 * J;
    // A full group of 8 frames is 32 J out_c / 288 J / 64 J bytes: whole 16-byte vectors.  Where the group's three
    // destinations are 16-byte aligned (they are whenever num_frames * J is even ... see vec_ok) the rows leave as
    // 16-byte stores: 4 x fewer store instructions than word stores -- this kernel is bound by instruction issue, not by
    // bytes (round 2: 38.8 us for 88 MB with word stores).
    const bool vec_ok = nf == P2R_DL_FRAMES && out_vec16 &&
                        ((reinterpret_cast<uintptr_t>(gj) | reinterpret_cast<uintptr_t>(gv) | reinterpret_cast<uintptr_t>(gm)) & 15) == 0;
    if (vec_ok) {
      const float4* sj4 = reinterpret_cast<const float4*>(s_out_j);
      const float4* sv4 = reinterpret_cast<const float4*>(s_out_v);
      const float4* sm4 = reinterpret_cast<const float4*>(s_mask);
      float4* gj4 = reinterpret_cast<float4*>(gj);
      float4* gv4 = reinterpret_cast<float4*>(gv);
      float4* gm4 = reinterpret_cast<float4*>(gm);
      for (int i = threadIdx.x; i < nf * rowo / 4; i += P2R_DL_THREADS) gj4[i] = sj4[i];
      for (int i = threadIdx.x; i < nf * rowl / 4; i += P2R_DL_THREADS) gv4[i] = sv4[i];
      for (int i = threadIdx.x; i < nf * J / 2; i += P2R_DL_THREADS) gm4[i] = sm4[i];
    } else {
      for (int i = threadIdx.x; i < nf * rowo; i += P2R_DL_THREADS) gj[i] = s_out_j[i];
      for (int i = threadIdx.x; i < nf * rowl; i += P2R_DL_THREADS) gv[i] = s_out_v[i];
      for (int i = threadIdx.x; i < nf * J; i += P2R_DL_THREADS) gm[i] = s_mask[i];
    }
    // the next iteration's first __syncthreads orders these reads of s_out_* / s_mask before they are rewritten
  }
  p2r_cp_async_wait<0>();
}

static int make_batch_launch(int variant, const float* joints, const float* votes, const long long* frame_start,
                             const int* sample_ids, const double* params, int b, int num_frames, int j, int out_channels,
                             float* input_joints, float* vote_label, long long* vote_label_mask, void* stream) {
  P2R_CHECK_ARG(b >= 0 && num_frames >= 0 && j > 0, "p2r_make_batch");
  P2R_CHECK_ARG(out_channels == 3 || out_channels == 4, "p2r_make_batch");
  P2R_CHECK_ARG(b <= 65535, "p2r_make_batch");
  P2R_CHECK_ARG(variant == 1 || variant == 2, "p2r_make_batch");
  if (b == 0 || num_frames == 0) return 0;
  if (variant == 2) {
    P2R_CHECK_ARG((reinterpret_cast<uintptr_t>(votes) & 3) == 0 && (reinterpret_cast<uintptr_t>(joints) & 3) == 0,
                  "p2r_make_batch (float arrays)");
    const int rowj = j * 3, rowv = j * 10;
    const size_t stage = (size_t)P2R_DL_FRAMES * (((rowv + 4 + 3) / 4) * 4 + ((rowj + 4 + 3) / 4) * 4) * sizeof(float);
    const size_t smem = (size_t)((P2R_DL_FRAMES * j * 2 + 3) / 4) * 4 * sizeof(float) + P2R_DL_STAGES * stage +
                        (size_t)P2R_DL_FRAMES * j * (out_channels + 9) * sizeof(float);
    P2R_CHECK_ARG(smem <= 200 * 1024, "p2r_make_batch");
    cudaError_t e = cudaFuncSetAttribute(make_batch_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { p2r_set_last_error("p2r_make_batch", (int)e); return (int)e; }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, make_batch_pipe_kernel, P2R_DL_THREADS, smem);
    if (e != cudaSuccess || per_sm < 1) { p2r_set_last_error("p2r_make_batch (occupancy query)", (int)e); return e ? (int)e : -1; }
    const int groups_per_item = p2r_ceil_div(num_frames, P2R_DL_FRAMES);
    const long long total = (long long)groups_per_item * b;
    const long long resident = (long long)P2R_SM_COUNT * per_sm;
    const int grid = (int)(total < resident ? total : resident);
    P2R_LAUNCH(make_batch_pipe_kernel, grid, P2R_DL_THREADS, smem, (cudaStream_t)stream, joints, votes, frame_start,
               sample_ids, params, num_frames, j, out_channels, groups_per_item, (int)total, input_joints, vote_label,
               vote_label_mask);
    P2R_RETURN_LAUNCH("p2r_make_batch");
  }
  const size_t smem = (size_t)P2R_DL_FRAMES * j * (sizeof(long long) + sizeof(float) * (3 + 10 + out_channels + 9));
  P2R_CHECK_ARG(smem <= 200 * 1024, "p2r_make_batch");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(make_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { p2r_set_last_error("p2r_make_batch", (int)e); return (int)e; }
  }
  dim3 grid(p2r_ceil_div(num_frames, P2R_DL_FRAMES), b);
  P2R_LAUNCH(make_batch_kernel, grid, P2R_DL_THREADS, smem, (cudaStream_t)stream, joints, votes, frame_start, sample_ids,
             params, num_frames, j, out_channels, input_joints, vote_label, vote_label_mask);
  P2R_RETURN_LAUNCH("p2r_make_batch");
}

extern "C" int p2r_make_batch(const float* joints, const float* votes, const long long* frame_start,
                              const int* sample_ids, const double* params, int b, int num_frames, int j,
                              int out_channels, float* input_joints, float* vote_label,
                              long long* vote_label_mask, void* stream) {
  static int variant = 0;
  if (variant == 0) {
    const char* e = getenv("P2R_MAKE_BATCH_VARIANT");
    variant = (e != nullptr && atoi(e) == 1) ? 1 : 2;   // 2 since round 2: 38.8 vs 86.2 us on a B200, bit-identical
  }
  return make_batch_launch(variant, joints, votes, frame_start, sample_ids, params, b, num_frames, j, out_channels,
                           input_joints, vote_label, vote_label_mask, stream);
}

extern "C" int p2r_make_batch_variant(int variant, const float* joints, const float* votes, const long long* frame_start,
                                      const int* sample_ids, const double* params, int b, int num_frames, int j,
                                      int out_channels, float* input_joints, float* vote_label,
                                      long long* vote_label_mask, void* stream) {
  return make_batch_launch(variant, joints, votes, frame_start, sample_ids, params, b, num_frames, j, out_channels,
                           input_joints, vote_label, vote_label_mask, stream);
}
