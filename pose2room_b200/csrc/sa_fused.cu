// sa_fused.cu -- the set-abstraction layer's "group -> shared MLP -> max over nsample" as ONE tcgen05 kernel (sm_100a).
//
// Replaces, for the live ProposalNet layer (PointnetSAModuleVotes, mlp = [256, 256, 256], bn = False, max pooling;
// /root/reference/external/pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py:220-256 with grouping_operation,
// pointnet2_utils.py:319-346): group_points -> Conv2d 1x1 + ReLU -> Conv2d 1x1 + ReLU -> max_pool2d over nsample.
// The (B, 256, npoint, nsample) grouped tensor and both MLP activations never exist in HBM (the first activation is
// written out only when the caller asks for it: training keeps it for the backward pass).
//
//   feats [B, N, 256] bf16 channel-last rows, idx [B, P, S] int32 (ball query), W1 / W2 [256, 256] bf16 (nn.Linear
//   layout = Conv2d 1x1 weight), b1 / b2 fp32  ->  out [B*P, 256] = max_s relu(W2 relu(W1 x_s + b1) + b2), arg-max u8.
//
// One persistent CTA per SM walks over tiles of 128 grouped rows (= 128 / S proposals):
//   warps 2-5 (128 threads)  gather the tile's 128 feature rows (512 B each, one row per warp instruction: coalesced
//                            16-byte cp.async) into the K-major, 128B-swizzled operand layout tcgen05 expects;
//   warp 0                   streams W1 / W2 through a 3-stage TMA ring (32 KB per 64-wide k-block; the 256 KB of
//                            weights stay L2-resident);
//   warp 1                   GEMM 1: D1[128 rows x 256] = X . W1^T (rows on TMEM lanes);
//   warps 2-5                epilogue 1: D1 + b1 -> ReLU -> bf16 -> shared memory, again as a K-major operand (a lane
//                            = a row writes its own 128-byte pieces: conflict-free) [+ 4 TMA stores of it, training];
//   warp 1                   GEMM 2, TRANSPOSED: D2^T[256 channels x 128 rows] = W2 . H^T (two M = 128 halves), so that
//                            the nsample rows of a proposal sit side by side in the COLUMNS of one TMEM lane;
//   warps 2-5                epilogue 2: a thread owns two output channels and takes the max / arg-max over each run
//                            of S columns in registers (no shuffles), + b2, ReLU, coalesced stores.
// The gather of tile j+1 overlaps GEMM 2 of tile j, GEMM 1 of tile j+1 overlaps epilogue 2 of tile j.
// TMEM: D1 = columns [0, 256), D2^T = [256, 512).  Shared memory: X 64 KB + H 64 KB + weight ring 96 KB.
#include "tcgen05.cuh"

#define SA_TILE 128
#define SA_C 256
#define SA_KB 4                     // 64-wide k-blocks per layer
#define SA_STAGES 3
#define SA_THREADS 192
#define SA_OPER_BYTES (SA_KB * SA_TILE * 128)      // one 128 x 256 bf16 operand, K-major: 64 KB
#define SA_STAGE_BYTES (SA_C * 128)                // one k-block of a weight matrix: 256 rows x 64 k: 32 KB
#define SA_SMEM_BYTES (2 * SA_OPER_BYTES + SA_STAGES * SA_STAGE_BYTES + SA_C * 4 + 128 + 1024 /*alignment slack*/)

__device__ __forceinline__ void sa_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(p2r_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void sa_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename OutT>
__global__ void __launch_bounds__(SA_THREADS, 1)
sa_fused_kernel(const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                const __grid_constant__ CUtensorMap map_h1, const __nv_bfloat16* __restrict__ feats,
                const int* __restrict__ idx, const float* __restrict__ b1, const float* __restrict__ b2, int n_points,
                int rows_per_batch, long long rows, int nsample, int save_h1, OutT* __restrict__ out,
                unsigned char* __restrict__ argmax, int ntiles) {
  extern __shared__ uint8_t sa_smem_raw[];
  uint8_t* sa_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sa_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* xb = sa_smem;                                   // gathered rows, 4 k-blocks x [128 rows][128 B]
  uint8_t* hb = sa_smem + SA_OPER_BYTES;                   // first activation, same layout
  uint8_t* ring = hb + SA_OPER_BYTES;                      // SA_STAGES x [256 rows][128 B]
  float* sb1 = reinterpret_cast<float*>(ring + SA_STAGES * SA_STAGE_BYTES);
  uint64_t* w_full = reinterpret_cast<uint64_t*>(sb1 + SA_C);
  uint64_t* w_empty = w_full + SA_STAGES;
  uint64_t* x_full = w_empty + SA_STAGES;
  uint64_t* d1_full = x_full + 1;
  uint64_t* h_full = d1_full + 1;
  uint64_t* d2_full = h_full + 1;
  uint64_t* d2_empty = d2_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SA_STAGES; ++s) {
      p2r_mbar_init(w_full + s, 1);
      p2r_mbar_init(w_empty + s, 1);
    }
    p2r_mbar_init(x_full, 4);      // one arrival per gather warp
    p2r_mbar_init(d1_full, 1);
    p2r_mbar_init(h_full, 4);
    p2r_mbar_init(d2_full, 1);
    p2r_mbar_init(d2_empty, 4);
    p2r_fence_mbar_init();
  }
  for (int c = threadIdx.x; c < SA_C; c += SA_THREADS) sb1[c] = b1 != nullptr ? __ldg(b1 + c) : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w1);
    tma_prefetch_desc(&map_w2);
    if (save_h1) tma_prefetch_desc(&map_h1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < my_tiles; ++j) {
        for (int i = 0; i < 2 * SA_KB; ++i, ++it) {
          const int s = it % SA_STAGES;
          p2r_mbar_wait(w_empty + s, ((uint32_t)(it / SA_STAGES) & 1u) ^ 1u);
          p2r_mbar_expect_tx(w_full + s, SA_STAGE_BYTES);
          tma_load_2d(ring + s * SA_STAGE_BYTES, i < SA_KB ? &map_w1 : &map_w2, w_full + s, (i % SA_KB) * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(128, 256, 0, 0);     // D1[128 rows, 256 ch] = X (A, K-major) . W1 (B, K-major)
      constexpr uint32_t idesc2 = make_idesc(128, 128, 0, 0);     // D2^T half[128 ch, 128 rows] = W2 half (A) . H (B)
      int it = 0;
      for (int j = 0; j < my_tiles; ++j) {
        const uint32_t par = (uint32_t)j & 1u;
        p2r_mbar_wait(x_full, par);
        tc_fence_after();
        for (int kb = 0; kb < SA_KB; ++kb, ++it) {
          const int s = it % SA_STAGES;
          p2r_mbar_wait(w_full + s, (uint32_t)(it / SA_STAGES) & 1u);
          tc_fence_after();
          const uint32_t a_addr = p2r_smem_u32(xb + kb * (SA_TILE * 128));
          const uint32_t b_addr = p2r_smem_u32(ring + s * SA_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, make_desc(a_addr + k * 32, 16, 1024), make_desc(b_addr + k * 32, 16, 1024), idesc1,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(w_empty + s);
        }
        umma_commit(d1_full);            // D1 complete (and every earlier MMA: X and, for j > 0, H are free again)
        p2r_mbar_wait(h_full, par);
        tc_fence_after();
        if (j > 0) {                     // epilogue 2 of the previous tile has drained D2
          p2r_mbar_wait(d2_empty, (uint32_t)(j - 1) & 1u);
          tc_fence_after();
        }
        for (int kb = 0; kb < SA_KB; ++kb, ++it) {
          const int s = it % SA_STAGES;
          p2r_mbar_wait(w_full + s, (uint32_t)(it / SA_STAGES) & 1u);
          tc_fence_after();
          const uint32_t a_addr = p2r_smem_u32(ring + s * SA_STAGE_BYTES);
          const uint32_t b_addr = p2r_smem_u32(hb + kb * (SA_TILE * 128));
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + 256u + (uint32_t)(h * 128), make_desc(a_addr + h * (128 * 128) + k * 32, 16, 1024),
                        make_desc(b_addr + k * 32, 16, 1024), idesc2, (kb | k) != 0 ? 1u : 0u);
          umma_commit(w_empty + s);
        }
        umma_commit(d2_full);
      }
    }
  } else {
    // ===================== gather + epilogues (warps 2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter of this warp = rows q*32 .. q*32+31 of the tile
    const int r_own = q * 32 + lane;
    const int log_s = 31 - __clz(nsample);  // nsample divides 128: a power of two
    const int ppt = SA_TILE >> log_s;       // proposals per tile
    const long long n_prop = rows >> log_s;
    const float bias2[2] = {b2 != nullptr ? __ldg(b2 + r_own) : 0.f, b2 != nullptr ? __ldg(b2 + 128 + r_own) : 0.f};

    auto gather = [&](int tile) {
      // lane l first fetches the source row of tile row q*32 + l, then the warp copies one 512-byte row per step
      const long long g = (long long)tile * SA_TILE + r_own;
      int src = 0;
      if (g < rows) src = (int)(g / rows_per_batch) * n_points + __ldg(idx + g);
#pragma unroll 4
      for (int i = 0; i < 32; ++i) {
        const int srow = __shfl_sync(0xffffffffu, src, i);
        const int r = q * 32 + i;
        const uint8_t* s = reinterpret_cast<const uint8_t*>(feats) + (size_t)srow * (SA_C * 2) + lane * 16;
        sa_cp_async16(xb + (lane >> 3) * (SA_TILE * 128) + r * 128 + (((lane & 7) ^ (r & 7)) << 4), s);
      }
      sa_cp_async_wait_all();
      p2r_fence_proxy_async();              // the tensor core reads shared memory through the async proxy
      __syncwarp();
      if (lane == 0) p2r_mbar_arrive(x_full);
    };

    if (my_tiles > 0) gather(blockIdx.x);
    for (int j = 0; j < my_tiles; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const uint32_t par = (uint32_t)j & 1u;
      // ---------------- epilogue 1: D1 (+ b1, ReLU) -> H in shared memory ----------------
      p2r_mbar_wait(d1_full, par);
      tc_fence_after();
      if (save_h1 && j > 0) {               // the bulk stores of the previous tile's H must have read it
        if (threadIdx.x == 64) tma_store_wait_read();
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
#pragma unroll 1
      for (int c0 = 0; c0 < SA_C; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        uint8_t* row = hb + (c0 >> 6) * (SA_TILE * 128) + r_own * 128;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[8 * p + e]) + sb1[c0 + 8 * p + e], 0.f);
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(f[6], f[7]);
          pk.x = *reinterpret_cast<uint32_t*>(&t0);
          pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2);
          pk.w = *reinterpret_cast<uint32_t*>(&t3);
          const int chunk = ((c0 & 63) >> 3) + p;
          *reinterpret_cast<uint4*>(row + ((chunk ^ (r_own & 7)) << 4)) = pk;
        }
      }
      tc_fence_before();
      p2r_fence_proxy_async();
      __syncwarp();
      if (lane == 0) p2r_mbar_arrive(h_full);
      if (save_h1) {                        // H[tile rows, 256] -> global (rows beyond `rows` are clipped by the map)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
#pragma unroll
          for (int kb = 0; kb < SA_KB; ++kb) tma_store_2d(&map_h1, hb + kb * (SA_TILE * 128), kb * 64, tile * SA_TILE);
        }
      }
      // ---------------- gather of the next tile (X is free: GEMM 1 of this tile has completed) ----------------
      if (j + 1 < my_tiles) gather(tile + gridDim.x);
      // ---------------- epilogue 2: D2^T -> max over each run of nsample columns, + b2, ReLU ----------------
      p2r_mbar_wait(d2_full, par);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int ch = h * 128 + r_own;
        const float bias = bias2[h];
        float best = 0.f;
        int bi = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < SA_TILE; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(h * 128 + c0), v);
          if (h == 1 && c0 + 32 == SA_TILE) {     // last read of D2: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) p2r_mbar_arrive(d2_empty);
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int n = c0 + e;
            const int sidx = n & (nsample - 1);
            const float x = __uint_as_float(v[e]);
            if (sidx == 0 || x > best) {          // first maximum wins, like F.max_pool2d / maxpool_rows
              best = x;
              bi = sidx;
            }
            if (sidx == nsample - 1) {
              const long long pg = (long long)tile * ppt + (n >> log_s);
              if (pg < n_prop) {
                const float y = fmaxf(best + bias, 0.f);
                if (sizeof(OutT) == 2) reinterpret_cast<__nv_bfloat16*>(out)[pg * SA_C + ch] = __float2bfloat16_rn(y);
                else reinterpret_cast<float*>(out)[pg * SA_C + ch] = y;
                if (argmax != nullptr) argmax[pg * SA_C + ch] = (unsigned char)bi;
              }
            }
          }
        }
      }
    }
    if (save_h1 && threadIdx.x == 64) tma_store_wait_all();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// C ABI: see include/p2r_b200.h.
extern "C" int p2r_sa_fused(const void* feats, const int* idx, const void* w1, const float* b1, const void* w2,
                            const float* b2, int b, int n, int p, int s, int c, void* out, int out_dtype,
                            unsigned char* argmax, void* h1, void* stream) {
  P2R_CHECK_ARG(b >= 0 && n > 0 && p >= 0 && s > 0, "p2r_sa_fused");
  P2R_CHECK_ARG(c == SA_C, "p2r_sa_fused (256-channel layers only)");
  P2R_CHECK_ARG(s <= SA_TILE && SA_TILE % s == 0, "p2r_sa_fused (nsample must divide 128)");
  P2R_CHECK_ARG(out_dtype == 0 || out_dtype == 1, "p2r_sa_fused");
  P2R_CHECK_ARG((long long)b * n < (1ll << 31), "p2r_sa_fused");
  P2R_CHECK_ARG((reinterpret_cast<uintptr_t>(feats) & 15) == 0, "p2r_sa_fused (feats must be 16-byte aligned)");
  const long long rows = (long long)b * p * s;
  if (rows == 0) return 0;
  const int ntiles = p2r_ceil_div(rows, SA_TILE);
  CUtensorMap m1, m2, mh;
  if (make_map(&m1, w1, SA_C, SA_C, SA_C, SA_C)) return -1;
  if (make_map(&m2, w2, SA_C, SA_C, SA_C, SA_C)) return -1;
  mh = m1;
  if (h1 != nullptr && make_map(&mh, h1, SA_C, rows, SA_C, SA_TILE)) return -1;
  const int grid = ntiles < P2R_SM_COUNT ? ntiles : P2R_SM_COUNT;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == 1) {
    auto kern = sa_fused_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SA_SMEM_BYTES);
    kern<<<grid, SA_THREADS, SA_SMEM_BYTES, st>>>(m1, m2, mh, (const __nv_bfloat16*)feats, idx, b1, b2, n, p * s, rows, s,
                                                  h1 != nullptr ? 1 : 0, (__nv_bfloat16*)out, argmax, ntiles);
  } else {
    auto kern = sa_fused_kernel<float>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SA_SMEM_BYTES);
    kern<<<grid, SA_THREADS, SA_SMEM_BYTES, st>>>(m1, m2, mh, (const __nv_bfloat16*)feats, idx, b1, b2, n, p * s, rows, s,
                                                  h1 != nullptr ? 1 : 0, (float*)out, argmax, ntiles);
  }
  P2R_RETURN_LAUNCH("p2r_sa_fused");
}
