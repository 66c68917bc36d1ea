// stream_bn.cu -- HBM-streaming BatchNorm kernels for the block activations of the ST-GCN backbone:
// [M, 64] bf16 matrices (M = B*T*V = 819 200 rows at the BASELINE shape, 105 MB each) that every BatchNorm2d /
// BatchNorm1d of st_gcn_block.tcn and the pos_embed / sk_feat MLPs (ref: stgcn_layers.py:402-414,436-438,
// stgcn.py:45-50) reads two to three times per direction.  These passes are pure HBM traffic, so the kernels are
// built around the copy engine instead of per-thread loads:
//
//   * a tile is 64 rows x 64 channels = 8 KB and CONTIGUOUS in memory, so one 1-D bulk TMA copy
//     (cp.async.bulk, SASS UBLKCP) per operand brings it into shared memory, completion on an mbarrier;
//   * one producer lane keeps a ring of STAGES tiles per operand in flight (full[] / empty[] mbarriers), 8 consumer
//     warps read 16-byte vectors from shared memory (conflict-free), do the arithmetic in fp32 and either accumulate
//     column sums in registers or write 16-byte vectors straight back to global memory;
//   * persistent CTAs (2 per SM), tiles interleaved across CTAs, per-channel coefficients in registers (a thread
//     always owns the same 8 channels).
//
// Four operations share the skeleton (same semantics as the generic kernels in dense_ops.cu they replace):
//   STATS_FWD : s1 = sum x, s2 = sum x^2
//   STATS_BWD : s1 = sum dz, s2 = sum dz * xhat            dz = dy masked by the ReLU (relu 1: y > 0, 2: x*sc+sh > 0)
//   AFFINE    : y = act(x * scale + shift (+ residual))
//   BWD_APPLY : dx = scale * (dz - s1/M - xhat * s2/M) (training) or scale * dz (eval);  dres = dz (optional)
#include "p2r_common.cuh"
#include "stream_bn.cuh"
#include <stdlib.h>

namespace {

constexpr int SB_ROWS = 64;                  // rows per tile
constexpr int SB_C = 64;                     // channels (bf16): one row = 128 bytes
constexpr int SB_TILE_BYTES = SB_ROWS * SB_C * 2;
constexpr int SB_CONSUMERS = 256;            // 8 consumer warps; warp 8 = producer
constexpr int SB_THREADS = SB_CONSUMERS + 32;

enum { STATS_FWD = 0, STATS_BWD = 1, AFFINE = 2, BWD_APPLY = 3 };

struct StreamArgs {
  const __nv_bfloat16* in[3];
  __nv_bfloat16* out[2];
  long long M;
  const float* scale;   // [64]
  const float* shift;
  const float* mean;
  const float* rstd;
  const double* s1;     // BWD_APPLY: column sums from STATS_BWD (NULL = eval mode)
  const double* s2;
  double inv_m;
  int relu;             // 0 none; 1 mask from in[2] (y > 0); 2 mask recomputed from x*scale+shift > 0; AFFINE: 0/1
  double* o1;           // STATS_*: outputs (atomically accumulated, zero-filled by the caller)
  double* o2;
  double* colsum;       // BWD_APPLY: optional sums of dx over the rows with the same (row % period): [period][64] doubles,
  int period;           //            atomically accumulated (the bias gradient of the layer that produced x: rows = joints)
  unsigned char* mask_out;        // AFFINE: optional ReLU bit mask of the output, [M, 8] bytes (bit i of byte cv = channel 8 cv + i)
  const unsigned char* mask_in;   // STATS_BWD / BWD_APPLY with relu == 3: that mask instead of re-reading y
  const double* sums64;           // BWD_APPLY: optional [2][64] doubles (the finished STATS_BWD sums) that block 0 also
  float* sums32;                  //            writes as float32 [2][64]: d beta / d gamma in the parameter's dtype
};

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  return v;
}

template <int MODE, int NIN, int STAGES>
__global__ void __launch_bounds__(SB_THREADS, 2) stream_bn_kernel(const StreamArgs a) {
  P2R_DYN_SMEM_ALIGNED(uint8_t, sb_smem, 128);
  constexpr int STAGE_BYTES = NIN * SB_TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sb_smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (a.M + SB_ROWS - 1) / SB_ROWS;

  // BWD_APPLY with a.colsum: per-(row % period, channel) sums of dx, accumulated with shared-memory float atomics
  // (row stride 72 floats + a per-lane rotation of the channel order make the 32 lanes of a warp hit 32 banks)
  constexpr int CS_LD = 72, CS_MAXP = 32;
  __shared__ float cs_tab[MODE == BWD_APPLY ? CS_MAXP * CS_LD : 1];
  // period == 1 (plain column sums of dx: the bias gradient of the conv in front of this BatchNorm) needs no table: every
  // thread sums its own 8 channels in registers (acc1, unused in this mode) and the CTA folds them like the STATS modes
  const bool do_cs1 = MODE == BWD_APPLY && a.colsum != nullptr && a.period == 1;
  const bool do_cs = MODE == BWD_APPLY && a.colsum != nullptr && !do_cs1;
  if (do_cs)
    for (int i = tid; i < CS_MAXP * CS_LD; i += SB_THREADS) cs_tab[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      p2r_mbar_init(full + s, 1);
      p2r_mbar_init(empty + s, SB_CONSUMERS / 32);
    }
    p2r_fence_mbar_init();
  }
  __syncthreads();
  if (MODE == BWD_APPLY && a.sums32 != nullptr && blockIdx.x == 0 && tid < 2 * SB_C) a.sums32[tid] = (float)a.sums64[tid];

  if (warp == SB_CONSUMERS / 32) {
    // ===================== producer: one lane streams tiles through the ring =====================
    if (lane == 0) {
      int it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        p2r_mbar_wait(empty + s, ((uint32_t)(it / STAGES) & 1u) ^ 1u);
        const long long r0 = tile * SB_ROWS;
        const uint32_t bytes = (uint32_t)min((long long)SB_ROWS, a.M - r0) * (SB_C * 2);
        p2r_mbar_expect_tx(full + s, bytes * NIN);
#pragma unroll
        for (int k = 0; k < NIN; ++k)
          p2r_bulk_g2s(sb_smem + s * STAGE_BYTES + k * SB_TILE_BYTES, a.in[k] + r0 * SB_C, bytes, full + s);
      }
    }
    return;
  }

  // ===================== consumers =====================
  const int cv = tid & 7;            // which 16-byte vector of the row: channels 8 cv .. 8 cv + 7
  const int rl = tid >> 3;           // row lane 0..31 (rows rl and rl + 32 of every tile)
  const int c0 = cv * 8;
  float sc[8], sh[8], A[8], Bc[8], D[8], acc1[8], acc2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = sh[i] = A[i] = Bc[i] = D[i] = acc1[i] = acc2[i] = 0.f;
    if (MODE == AFFINE || ((MODE == STATS_BWD || MODE == BWD_APPLY) && a.relu == 2) || MODE == BWD_APPLY) {
      if (a.scale) sc[i] = __ldg(a.scale + c0 + i);
      if (a.shift) sh[i] = __ldg(a.shift + c0 + i);
    }
    if (MODE == BWD_APPLY) {
      // dx = scale * (dz - k1 - xhat * k2),  xhat = (x - mean) * rstd   ==>   dx = dz * A + x * B + D
      A[i] = sc[i];
      if (a.s1) {
        const float k1 = (float)(a.s1[c0 + i] * a.inv_m), k2 = (float)(a.s2[c0 + i] * a.inv_m);
        const float mu = __ldg(a.mean + c0 + i), rs = __ldg(a.rstd + c0 + i);
        Bc[i] = -rs * k2 * sc[i];
        D[i] = (-k1 + mu * rs * k2) * sc[i];
      }
    }
  }

  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % STAGES;
    p2r_mbar_wait(full + s, (uint32_t)(it / STAGES) & 1u);
    const long long r0 = tile * SB_ROWS;
    const int rows_here = (int)min((long long)SB_ROWS, a.M - r0);
    const uint8_t* st = sb_smem + s * STAGE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = rl + 32 * h;
      if (r < rows_here) {
        const int off = r * (SB_C * 2) + cv * 16;
        float p[8], q[8], w[8];
        unpack8(*reinterpret_cast<const uint4*>(st + off), p);
        if (NIN >= 2) unpack8(*reinterpret_cast<const uint4*>(st + SB_TILE_BYTES + off), q);
        if (NIN >= 3) unpack8(*reinterpret_cast<const uint4*>(st + 2 * SB_TILE_BYTES + off), w);
        const size_t go = (size_t)(r0 + r) * SB_C + c0;
        if (MODE == STATS_FWD) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc1[i] += p[i]; acc2[i] = fmaf(p[i], p[i], acc2[i]); }
        } else if (MODE == STATS_BWD) {
          // p = dy, q = x (NIN >= 2), w = y (NIN == 3, relu == 1); relu == 3: one mask byte per thread instead of y
          const unsigned mb = (NIN == 2 && a.relu == 3) ? (unsigned)__ldg(a.mask_in + (size_t)(r0 + r) * 8 + cv) : 0xffu;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float dz = p[i];
            if (NIN >= 3 && !(w[i] > 0.f)) dz = 0.f;
            if (NIN == 2 && a.relu == 2 && !(fmaf(q[i], sc[i], sh[i]) > 0.f)) dz = 0.f;
            if (NIN == 2 && !((mb >> i) & 1u)) dz = 0.f;
            acc1[i] += dz;
            if (NIN >= 2) acc2[i] = fmaf(dz, q[i], acc2[i]);
          }
        } else if (MODE == AFFINE) {
          // p = x, q = residual (NIN == 2)
          float y[8];
          unsigned bits = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t = fmaf(p[i], sc[i], sh[i]);
            if (NIN >= 2) t += q[i];
            y[i] = a.relu ? fmaxf(t, 0.f) : t;
          }
          const uint4 packed = pack8(y);
          *reinterpret_cast<uint4*>(a.out[0] + go) = packed;
          if (a.mask_out != nullptr) {      // the mask of the STORED (bf16-rounded) output, as the backward would see it
            float yr[8];
            unpack8(packed, yr);
#pragma unroll
            for (int i = 0; i < 8; ++i) bits |= (yr[i] > 0.f ? 1u : 0u) << i;
            a.mask_out[(size_t)(r0 + r) * 8 + cv] = (unsigned char)bits;
          }
        } else {
          // BWD_APPLY: p = dy, q = x, w = y (NIN == 3, relu == 1); relu == 3: mask byte
          const unsigned mb = (NIN == 2 && a.relu == 3) ? (unsigned)__ldg(a.mask_in + (size_t)(r0 + r) * 8 + cv) : 0xffu;
          float g[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float dz = p[i];
            if (NIN >= 3 && !(w[i] > 0.f)) dz = 0.f;
            if (NIN == 2 && a.relu == 2 && !(fmaf(q[i], sc[i], sh[i]) > 0.f)) dz = 0.f;
            if (NIN == 2 && !((mb >> i) & 1u)) dz = 0.f;
            p[i] = dz;
            g[i] = fmaf(dz, A[i], fmaf(q[i], Bc[i], D[i]));
          }
          if (a.out[1]) *reinterpret_cast<uint4*>(a.out[1] + go) = pack8(p);
          const uint4 gp = pack8(g);
          *reinterpret_cast<uint4*>(a.out[0] + go) = gp;
          if (do_cs1) {     // sums of the STORED (bf16) dx, like a separate pass over dx would see them
            float gr[8];
            unpack8(gp, gr);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc1[i] += gr[i];
          }
          if (do_cs) {      // sums of the STORED (bf16) dx, like a separate pass over dx would see them
            float gr[8];
            unpack8(gp, gr);
            float* row = cs_tab + (int)((r0 + r) % a.period) * CS_LD + c0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int ii = (i + lane) & 7;
              atomicAdd(row + ii, gr[ii]);
            }
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) p2r_mbar_arrive(empty + s);   // this warp no longer reads stage s
  }

  if (MODE == BWD_APPLY) {
    if (do_cs) {
      P2R_NAMED_BARRIER_SYNC_1_256();   // every consumer has added its last row
      for (int i = tid; i < a.period * SB_C; i += SB_CONSUMERS) {
        const float v = cs_tab[(i / SB_C) * CS_LD + (i % SB_C)];
        if (v != 0.f) atomicAdd(a.colsum + i, (double)v);
      }
    }
  }
  if (MODE == BWD_APPLY && do_cs1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc1[i] += __shfl_xor_sync(0xffffffffu, acc1[i], 8);
      acc1[i] += __shfl_xor_sync(0xffffffffu, acc1[i], 16);
    }
    P2R_NAMED_BARRIER_SYNC_1_256();   // every consumer is past its last tile: the ring is free
    float* red = reinterpret_cast<float*>(sb_smem);  // [8 warps][64 channels]
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp * 64 + c0 + i] = acc1[i];
    }
    P2R_NAMED_BARRIER_SYNC_1_256();
    if (tid < 64) {
      double t = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += (double)red[w8 * 64 + tid];
      atomicAdd(a.colsum + tid, t);
    }
  }
  if (MODE == STATS_FWD || MODE == STATS_BWD) {
    // lanes l, l^8, l^16, l^24 own the same channels: fold them, then combine the 8 warps through shared memory
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc1[i] += __shfl_xor_sync(0xffffffffu, acc1[i], 8);
      acc1[i] += __shfl_xor_sync(0xffffffffu, acc1[i], 16);
      acc2[i] += __shfl_xor_sync(0xffffffffu, acc2[i], 8);
      acc2[i] += __shfl_xor_sync(0xffffffffu, acc2[i], 16);
    }
    P2R_NAMED_BARRIER_SYNC_1_256();   // every consumer is past its last tile: the ring is free
    float* red = reinterpret_cast<float*>(sb_smem);  // [2][8 warps][64 channels]
    if (lane < 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[(0 * 8 + warp) * 64 + c0 + i] = acc1[i];
        red[(1 * 8 + warp) * 64 + c0 + i] = acc2[i];
      }
    }
    P2R_NAMED_BARRIER_SYNC_1_256();
    if (tid < 128) {
      const int which = tid >> 6, c = tid & 63;
      double t = 0.0;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += (double)red[(which * 8 + w8) * 64 + c];
      if (which == 0) atomicAdd(a.o1 + c, t);
      else if (a.o2 != nullptr) {
        if (MODE == STATS_BWD && a.mean != nullptr) {
          double t1 = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) t1 += (double)red[w8 * 64 + c];
          t = (double)__ldg(a.rstd + c) * (t - (double)__ldg(a.mean + c) * t1);
        }
        atomicAdd(a.o2 + c, t);
      }
    }
  }
}

// Column sums of a matrix whose rows cycle through `period` classes (the graph convolution's bias gradient: dG viewed as
// [M * V, 64], a row's class = its joint): out[row % period][channel] += x.  Same producer / consumer ring, but a tile is
// a whole number of periods (`tile_rows` = period * floor(64 / period) rows), so the class of a thread's rows never
// changes and the sums live in registers; the generic colsum_wide_kernel (per-thread loads) ran at 1.7 TB/s.
__global__ void __launch_bounds__(SB_THREADS, 2)
stream_colsum_period_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int period, int tile_rows,
                            double* __restrict__ out) {
  P2R_DYN_SMEM_ALIGNED(uint8_t, sb_smem, 128);
  constexpr int STAGES = 8;
  const int tile_bytes = tile_rows * (SB_C * 2);
  uint64_t* full = reinterpret_cast<uint64_t*>(sb_smem + STAGES * SB_TILE_BYTES);
  uint64_t* empty = full + STAGES;
  __shared__ float cs_tab[32 * SB_C];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (rows + tile_rows - 1) / tile_rows;
  for (int i = tid; i < 32 * SB_C; i += SB_THREADS) cs_tab[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      p2r_mbar_init(full + s, 1);
      p2r_mbar_init(empty + s, SB_CONSUMERS / 32);
    }
    p2r_fence_mbar_init();
  }
  __syncthreads();
  if (warp == SB_CONSUMERS / 32) {
    if (lane == 0) {
      int it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        p2r_mbar_wait(empty + s, ((uint32_t)(it / STAGES) & 1u) ^ 1u);
        const long long r0 = tile * tile_rows;
        const uint32_t bytes = (uint32_t)min((long long)tile_rows, rows - r0) * (SB_C * 2);
        p2r_mbar_expect_tx(full + s, bytes);
        p2r_bulk_g2s(sb_smem + s * SB_TILE_BYTES, x + r0 * SB_C, bytes, full + s);
      }
    }
    return;
  }
  const int cv = tid & 7, rl = tid >> 3, c0 = cv * 8;
  float acc[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[h][i] = 0.f;
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % STAGES;
    p2r_mbar_wait(full + s, (uint32_t)(it / STAGES) & 1u);
    const int rows_here = (int)min((long long)tile_rows, rows - tile * tile_rows);
    const uint8_t* st = sb_smem + s * SB_TILE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = rl + 32 * h;
      if (r < rows_here) {
        float p[8];
        unpack8(*reinterpret_cast<const uint4*>(st + r * (SB_C * 2) + cv * 16), p);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[h][i] += p[i];
      }
    }
    __syncwarp();
    if (lane == 0) p2r_mbar_arrive(empty + s);
  }
  (void)tile_bytes;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = rl + 32 * h;
    if (r < tile_rows) {
      float* row = cs_tab + (r % period) * SB_C + c0;
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(row + i, acc[h][i]);
    }
  }
  P2R_NAMED_BARRIER_SYNC_1_256();
  for (int i = tid; i < period * SB_C; i += SB_CONSUMERS) {
    const float v = cs_tab[i];
    if (v != 0.f) atomicAdd(out + i, (double)v);
  }
}

int stream_ctas_per_sm() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("P2R_STREAM_CTAS_PER_SM");
    v = e ? atoi(e) : 2;
    if (v < 1) v = 1;
    if (v > 2) v = 2;
  }
  return v;
}

template <int MODE, int NIN>
int launch_stream(const StreamArgs& a, cudaStream_t st, const char* where) {
  constexpr int STAGES = NIN == 1 ? 8 : (NIN == 2 ? 6 : 4);     // 64 - 96 KB of tiles in flight per CTA
  constexpr int SMEM = STAGES * NIN * SB_TILE_BYTES + 2 * STAGES * 8;
  const long long ntiles = (a.M + SB_ROWS - 1) / SB_ROWS;
  const int grid = (int)min(ntiles, (long long)P2R_SM_COUNT * stream_ctas_per_sm());
  auto kern = stream_bn_kernel<MODE, NIN, STAGES>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  P2R_LAUNCH(kern, grid, SB_THREADS, SMEM, st, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) p2r_set_last_error(where, (int)e);
  return (int)e;
}

bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool p2r_stream_bn_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("P2R_STREAM_BN");
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

extern "C" int p2r_stream_bn_supported(int dtype, long long M, int C) {
  return (p2r_stream_bn_enabled() && dtype == 1 && C == SB_C && M >= 4096) ? 1 : 0;
}

bool p2r_stream_bn_ok(int dtype, long long M, int C, const void* p0, const void* p1, const void* p2, const void* p3,
                      const void* p4) {
  return p2r_stream_bn_enabled() && dtype == 1 && C == SB_C && M >= 4096 && aligned16(p0) && aligned16(p1) &&
         aligned16(p2) && aligned16(p3) && aligned16(p4);
}

// out: double[period * 64], zero-filled by the caller.  rows must be a multiple of period, 2 <= period <= 32.
int p2r_stream_colsum_period(const void* x, long long rows, int period, double* out, cudaStream_t st) {
  const int tile_rows = period * (SB_ROWS / period);
  constexpr int SMEM = 8 * SB_TILE_BYTES + 2 * 8 * 8;
  const long long ntiles = (rows + tile_rows - 1) / tile_rows;
  const int grid = (int)min(ntiles, (long long)P2R_SM_COUNT * stream_ctas_per_sm());
  cudaFuncSetAttribute(stream_colsum_period_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  P2R_LAUNCH(stream_colsum_period_kernel, grid, SB_THREADS, SMEM, st, (const __nv_bfloat16*)x, rows, period, tile_rows, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) p2r_set_last_error("p2r_col_sum_wide", (int)e);
  return (int)e;
}

int p2r_stream_col_stats(const void* x, long long M, double* s1, double* s2, cudaStream_t st) {
  StreamArgs a = {};
  a.in[0] = (const __nv_bfloat16*)x;
  a.M = M;
  a.o1 = s1;
  a.o2 = s2;
  return launch_stream<STATS_FWD, 1>(a, st, "p2r_col_stats");
}

int p2r_stream_col_bwd_stats(const void* dy, const void* x, const void* y, long long M, const float* mean,
                             const float* rstd, int relu, double* s1, double* s2, const float* scale,
                             const float* shift, cudaStream_t st) {
  StreamArgs a = {};
  a.in[0] = (const __nv_bfloat16*)dy;
  a.in[1] = (const __nv_bfloat16*)x;
  a.in[2] = relu == 3 ? nullptr : (const __nv_bfloat16*)y;
  a.mask_in = relu == 3 ? (const unsigned char*)y : nullptr;
  a.M = M;
  a.mean = mean;
  a.rstd = rstd;
  a.scale = scale;
  a.shift = shift;
  a.relu = relu;
  a.o1 = s1;
  a.o2 = s2;
  if (x == nullptr) return launch_stream<STATS_BWD, 1>(a, st, "p2r_col_bwd_stats");   // relu == 0, only s1 (bias grad)
  if (relu == 1) return launch_stream<STATS_BWD, 3>(a, st, "p2r_col_bwd_stats");
  return launch_stream<STATS_BWD, 2>(a, st, "p2r_col_bwd_stats");
}

int p2r_stream_affine_act(const void* x, long long M, const float* scale, const float* shift, const void* residual,
                          int relu, void* y, unsigned char* relu_mask, cudaStream_t st) {
  StreamArgs a = {};
  a.mask_out = relu_mask;
  a.in[0] = (const __nv_bfloat16*)x;
  a.in[1] = (const __nv_bfloat16*)residual;
  a.out[0] = (__nv_bfloat16*)y;
  a.M = M;
  a.scale = scale;
  a.shift = shift;
  a.relu = relu;
  if (residual) return launch_stream<AFFINE, 2>(a, st, "p2r_affine_act");
  return launch_stream<AFFINE, 1>(a, st, "p2r_affine_act");
}

int p2r_stream_bn_bwd_apply(const void* dy, const void* x, const void* y, long long M, const float* mean,
                            const float* rstd, const float* scale, const double* s1, const double* s2, int relu,
                            void* dx, void* dres, const float* shift, double* colsum, int period, cudaStream_t st,
                            const double* sums64, float* sums32) {
  StreamArgs a = {};
  a.sums64 = sums64;
  a.sums32 = sums64 != nullptr ? sums32 : nullptr;
  a.colsum = (colsum != nullptr && period >= 1 && period <= 32) ? colsum : nullptr;
  a.period = period;
  a.in[0] = (const __nv_bfloat16*)dy;
  a.in[1] = (const __nv_bfloat16*)x;
  a.in[2] = relu == 3 ? nullptr : (const __nv_bfloat16*)y;
  a.mask_in = relu == 3 ? (const unsigned char*)y : nullptr;
  a.out[0] = (__nv_bfloat16*)dx;
  a.out[1] = (__nv_bfloat16*)dres;
  a.M = M;
  a.mean = mean;
  a.rstd = rstd;
  a.scale = scale;
  a.shift = shift;
  a.s1 = s1;
  a.s2 = s2;
  a.inv_m = 1.0 / (double)M;
  a.relu = relu;
  if (relu == 1) return launch_stream<BWD_APPLY, 3>(a, st, "p2r_bn_bwd_apply");
  return launch_stream<BWD_APPLY, 2>(a, st, "p2r_bn_bwd_apply");
}
