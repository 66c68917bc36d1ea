// gemm_sm100.cu -- bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared
// memory -> tcgen05.mma (accumulator in TMEM) -> tcgen05.ld epilogue.  This is the kernel behind every dense
// per-point layer of the P2RNet hot path in throughput mode, above all the fused graph-convolution GEMM
// (M = B*T frames, N = K = V*64) that replaces the reference's 1x1 conv 64->704 + einsum 'nkctv,kvw->nctw'
// (/root/reference/models/p2rnet/modules/stgcn_layers.py:58-67).
//
//   C[M,N] (+)= A . B^T,   fp32 accumulate, C in bf16 or fp32, optional bias[N], optional ReLU.
//   A is given either K-major ([M,K] row-major) or MN-major ([K,M] row-major),
//   B is given either K-major ([N,K] row-major, the nn.Linear weight layout) or MN-major ([K,N] row-major);
//   with these four combinations forward (x.W^T), input gradient (dy.W) and weight gradient (dy^T.x) all run
//   on the same kernel without materialising a transpose.
//
// Structure (one 128 x BLOCK_N output tile per CTA, 192 threads):
//   warp 0   : TMA producer  -- one elected lane issues the bulk-tensor loads of the next k-block into a
//              STAGES-deep ring, completion signalled on full[] mbarriers (expect_tx)
//   warp 1   : TMEM allocator + MMA issuer -- one elected lane issues 4 x tcgen05.mma (K=16 each) per
//              64-wide k-block, tcgen05.commit releases the smem stage (empty[]) and finally signals tmem_full
//   warps 2-5: epilogue -- tcgen05.ld 32 lanes x 32 columns at a time, bias/ReLU/convert, 16-byte global stores
//              (or fp32 atomics for split-K)
// Two CTAs are co-resident per SM (smem <= 110 KB, TMEM <= 256 columns each) so one CTA's epilogue overlaps
// the other's main loop.
#include "tcgen05.cuh"

#define GEMM_BLOCK_M 128
#define GEMM_BLOCK_K 64
#define GEMM_THREADS 192


// Temporal-tap addressing (the (3x1) temporal convolution of st_gcn_block.tcn as an implicit GEMM, no unfold):
//   TAP = 1: A is a 3-D tensor (C, rows_per_sample, samples); k-block i belongs to tap i / kb_per_tap and reads the
//            A rows shifted by tap_shift0 + tap * tap_shift_step; rows outside a sample are zero-filled by TMA.
//   TAP = 2: B (MN-major) is that 3-D tensor; the 64-wide n-block n selects tap n / C and channel offset n % C,
//            and the reduction rows (k) are shifted per tap.
struct TapArgs {
  int kb_per_tap, rows_per_sample, channels, shift0, shift_step;
};

// TAP == 3 (temporal conv forward / input gradient with ONE halo tile per 128 output rows instead of three shifted 128-row
// views: 23 KB instead of 48 KB through L2 -> SM per tile, and the 24 KB of weights loaded once per CTA instead of once
// per tile -- the three-view kernels moved 472 MB per launch through the crossbar for 165 MB of DRAM traffic and were
// bound by it).  Shared memory: [weights 3 x 8 KB][HALO_STAGES x halo tile][staging][bias][barriers].
#define HALO_STAGES 3
#define HALO_W_BYTES (3 * 8192)
__host__ __device__ inline int halo_tile_bytes(int v) { return ((GEMM_BLOCK_M + 2 * v) * 128 + 1023) & ~1023; }
// TAP == 4 (temporal conv weight gradient): the three-view loads of TAP == 2 (one N = 192 MMA per 16 reduction rows; three
// N = 64 MMAs on row-offset views of one halo tile were tried and are SLOWER, 93 vs 74 us: issuing a tcgen05.mma costs
// ~60 ns whatever its size), but only the live half of the dy tile is staged -- Co = 64 of the M = 128 operand rows; the
// other 64-row block of every stage's descriptor points at one shared zero block -- so a stage is 32 KB instead of 40 KB
// and six of them fit instead of four: the kernel is bound by how many reduction rows are in flight.
#define DW4_A_BYTES (GEMM_BLOCK_K * 128)
#define DW4_STAGE_BYTES (DW4_A_BYTES + 3 * GEMM_BLOCK_K * 128)

// Optional extras of one launch (all NULL / 0 = plain dense GEMM):
//   kb_list   : block-sparse reduction.  Row n_tile of an int table [tiles_n][kb_stride]: entry 0 = number of 64-wide
//               k-blocks this n-tile visits, entries 1.. = their indices (ascending).  The graph-convolution weight
//               W_eff = sum_k W_k (x) A_k is zero in every 64x64 block (w,v) whose joints are further apart than the
//               adjacency's max hop, so only the listed k-blocks are loaded and multiplied.
//   tile_mask : [tiles_m][tiles_n] bytes, 0 = this output tile is structurally zero: the CTA exits (C untouched).
//   stats     : fused BatchNorm statistics of the OUTPUT: per channel c = column % 64, sum and sum of squares of the
//               values as stored (after bias / rounding to the output type), accumulated with double atomics into
//               stats[copy][0][c] / stats[copy][1][c], copy = CTA index % stat_copies (spreads the atomic traffic).
// Diagnostic: per-tile timestamps (globaltimer, ns) of the three roles of ONE CTA of the halo temporal-conv kernels
// (TAP == 3): trace[((role * 64 + tile) * 8 + event)], role 0 producer / 1 MMA issuer / 2 first epilogue warp.  NULL (the
// default) = off; set with p2r_debug_tconv_trace().
__device__ long long* g_tconv_trace = nullptr;
__device__ __forceinline__ void trace_mark(long long* tr, int role, int tile, int ev) {
  if (tr != nullptr && tile < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tr[(role * 64 + tile) * 8 + ev] = (long long)t;
  }
}

struct GemmExtra {
  const int* kb_list;
  int kb_stride;
  const unsigned char* tile_mask;
  double* stats;
  int stat_copies;
};

// Sum over the 32 lanes of a warp of 32 per-lane values, transposed: on return lane l holds sum_lanes v[l].
// 31 shuffles (16 + 8 + 4 + 2 + 1) instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}


template <int BLOCK_N, int STAGES_OVERRIDE = 0, bool TS = false>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = STAGES_OVERRIDE > 0 ? STAGES_OVERRIDE : ((BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 3 : 4));
  static constexpr int ACC_STAGES = BLOCK_N <= 128 ? 2 : 1;   // double-buffered accumulator when 2 x 2 CTAs fit in TMEM
  static constexpr int ACC_COLS = ACC_STAGES * BLOCK_N;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : (ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : 256));
  // TS (TMA-store epilogue): a 32 x 64 bf16 staging tile per epilogue warp + the tile's bias, double-buffered
  static constexpr int STAGING_BYTES = TS ? 4 * 32 * 128 : 0;
  static constexpr int BIAS_BYTES = TS ? 2 * BLOCK_N * 4 : 0;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + BIAS_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// TS = true: the epilogue stages bf16 tiles in shared memory and writes them with bulk-tensor (TMA) stores through
// tma_c (box {64 columns, 32 rows}); see the CTA-pair kernel below for why.  Needs bf16 C, BLOCK_N % 64 == 0.
template <int BLOCK_N, bool A_MN, bool B_MN, typename OutT, bool ATOMIC, int TAP = 0, int NSTAGE = 0, bool STATS = false,
          bool TS = false>
__global__ void __launch_bounds__(TAP == 3 ? GEMM_THREADS + 128 : GEMM_THREADS)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c,
                 OutT* __restrict__ C, int ldc, int M, int N, int K, const float* __restrict__ bias, int relu,
                 int kblocks_per_split, const TapArgs tap, int tiles_per_cta, const GemmExtra ex) {
  using S = GemmSmem<BLOCK_N, NSTAGE, TS>;
  if (ex.tile_mask != nullptr && ex.tile_mask[blockIdx.y * gridDim.x + blockIdx.x] == 0) return;   // whole CTA, uniform
  constexpr int ACC = S::ACC_STAGES;   // accumulator buffers in TMEM (2: the epilogue of tile t overlaps the MMAs of t+1)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int halo_v = tap.shift_step < 0 ? -tap.shift_step : tap.shift_step;  // (TAP 3) rows between two taps
  const int halo_bytes = halo_tile_bytes(halo_v);
  constexpr int stage4_bytes = DW4_STAGE_BYTES;                              // (TAP 4) live dy half + three x taps
  uint8_t* staging = smem + (TAP == 3 ? HALO_W_BYTES + HALO_STAGES * halo_bytes
                                      : (TAP == 4 ? S::STAGES * stage4_bytes + DW4_A_BYTES : S::STAGES * S::STAGE_BYTES));   // (TS) 1024-byte aligned
  float* sbias = reinterpret_cast<float*>(staging + S::STAGING_BYTES);       // (TS) [2][BLOCK_N]
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + S::STAGING_BYTES + S::BIAS_BYTES);
  uint64_t* empty = full + S::STAGES;
  uint64_t* tmem_full = empty + S::STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int total_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int kb0 = blockIdx.z * kblocks_per_split;
  const int kb1 = min(total_kb, kb0 + kblocks_per_split);
  const int* kbl = ex.kb_list != nullptr ? ex.kb_list + (size_t)blockIdx.x * ex.kb_stride + 1 : nullptr;
  const int nkb = TAP == 3 ? 3 : (kbl != nullptr ? __ldg(kbl - 1) : kb1 - kb0);
  const int tiles_m = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tile0 = blockIdx.y * tiles_per_cta;
  const int ntiles = min(tiles_per_cta, tiles_m - tile0);   // m-tiles this CTA walks through

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      p2r_mbar_init(full + s, 1);
      p2r_mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      p2r_mbar_init(tmem_full + a, 1);
      p2r_mbar_init(tmem_empty + a, TAP == 3 ? 8 : 4);   // one arrival per epilogue warp
    }
    p2r_fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (TS) tma_prefetch_desc(&tma_c);
  }
  if (TAP == 4) {   // rows 64..127 of the M = 128 dy operand do not exist (Co = 64): one zero block after the ring
    uint4* z = reinterpret_cast<uint4*>(smem + S::STAGES * stage4_bytes);
    for (int i = threadIdx.x; i < DW4_A_BYTES / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    p2r_fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (TAP == 3) {
      if (lane == 0) {
        p2r_mbar_expect_tx(full + HALO_STAGES, HALO_W_BYTES);           // the weights of the three taps, once
#pragma unroll
        for (int tp = 0; tp < 3; ++tp) {
          if (!B_MN) tma_load_2d(smem + tp * 8192, &tma_b, full + HALO_STAGES, tp * 64, n0);    // box {64 k, 64 n}
          else tma_load_2d(smem + tp * 8192, &tma_b, full + HALO_STAGES, n0, tp * 64);           // box {64 n, 64 k}
        }
        long long* tr = (blockIdx.y == 1 && blockIdx.x == 0) ? g_tconv_trace : nullptr;
        for (int t = 0; t < ntiles; ++t) {
          const int m0 = (tile0 + t) * GEMM_BLOCK_M;
          const int s = t % HALO_STAGES;
          p2r_mbar_wait(empty + s, ((uint32_t)(t / HALO_STAGES) & 1u) ^ 1u);
          trace_mark(tr, 0, t, 0);
          p2r_mbar_expect_tx(full + s, (uint32_t)(GEMM_BLOCK_M + 2 * halo_v) * 128u);
          tma_load_3d(smem + HALO_W_BYTES + s * halo_bytes, &tma_a, full + s, 0, (m0 % tap.rows_per_sample) - halo_v,
                      m0 / tap.rows_per_sample);                          // box {64 c, 128 + 2 V rows, 1 sample}
          trace_mark(tr, 0, t, 1);
        }
      }
    } else if (TAP == 4) {
      if (lane == 0) {
        for (int i = 0; i < nkb; ++i) {
          const int s = i % S::STAGES;
          p2r_mbar_wait(empty + s, ((uint32_t)(i / S::STAGES) & 1u) ^ 1u);
          uint8_t* a_dst = smem + s * stage4_bytes;
          const int k0 = (kb0 + i) * GEMM_BLOCK_K;
          p2r_mbar_expect_tx(full + s, (uint32_t)stage4_bytes);
          tma_load_2d(a_dst, &tma_a, full + s, 0, k0);                                            // box {64 co, 64 k}
#pragma unroll
          for (int tp = 0; tp < 3; ++tp)                                                          // boxes {64 ci, 64 rows, 1}
            tma_load_3d(a_dst + DW4_A_BYTES + tp * (GEMM_BLOCK_K * 128), &tma_b, full + s, 0,
                        (k0 % tap.rows_per_sample) + tap.shift0 + tp * tap.shift_step, k0 / tap.rows_per_sample);
        }
      }
    } else if (lane == 0) {
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int m0 = (tile0 + t) * GEMM_BLOCK_M;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (uint32_t)(it / S::STAGES) & 1u;
          p2r_mbar_wait(empty + s, ph ^ 1u);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          p2r_mbar_expect_tx(full + s, S::STAGE_BYTES);
          const int kbi = kbl != nullptr ? __ldg(kbl + i) : kb0 + i;
          const int k0 = kbi * GEMM_BLOCK_K;
          if (TAP == 1) {
            const int tp = kbi / tap.kb_per_tap;
            const int kc = (kbi % tap.kb_per_tap) * GEMM_BLOCK_K;
            tma_load_3d(a_dst, &tma_a, full + s, kc, (m0 % tap.rows_per_sample) + tap.shift0 + tp * tap.shift_step,
                        m0 / tap.rows_per_sample);                        // box {64 c, 128 rows, 1 sample}
          } else if (!A_MN) {
            tma_load_2d(a_dst, &tma_a, full + s, k0, m0);                 // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int h = 0; h < GEMM_BLOCK_M / 64; ++h)                   // boxes {64 m, 64 k}
              tma_load_2d(a_dst + h * (GEMM_BLOCK_K * 128), &tma_a, full + s, m0 + h * 64, k0);
          }
          if (TAP == 2) {
#pragma unroll
            for (int h = 0; h < BLOCK_N / 64; ++h) {                      // boxes {64 c, 64 rows, 1 sample}
              const int n = n0 + h * 64;
              const int tp = n / tap.channels;
              tma_load_3d(b_dst + h * (GEMM_BLOCK_K * 128), &tma_b, full + s, n % tap.channels,
                          (k0 % tap.rows_per_sample) + tap.shift0 + tp * tap.shift_step, k0 / tap.rows_per_sample);
            }
          } else if (!B_MN) {
            tma_load_2d(b_dst, &tma_b, full + s, k0, n0);                 // box {64 k, BLOCK_N n}
          } else {
#pragma unroll
            for (int h = 0; h < BLOCK_N / 64; ++h)                        // boxes {64 n, 64 k}
              tma_load_2d(b_dst + h * (GEMM_BLOCK_K * 128), &tma_b, full + s, n0 + h * 64, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (TAP == 3) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(GEMM_BLOCK_M, BLOCK_N, 0, B_MN ? 1 : 0);
        p2r_mbar_wait(full + HALO_STAGES, 0u);
        long long* tr = (blockIdx.y == 1 && blockIdx.x == 0) ? g_tconv_trace : nullptr;
        for (int t = 0; t < ntiles; ++t) {
          const int as = t % ACC;
          if (t >= ACC) {
            p2r_mbar_wait(tmem_empty + as, (uint32_t)((t / ACC) - 1) & 1u);
            tc_fence_after();
          }
          trace_mark(tr, 1, t, 0);
          const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
          const int s = t % HALO_STAGES;
          p2r_mbar_wait(full + s, (uint32_t)(t / HALO_STAGES) & 1u);
          tc_fence_after();
          trace_mark(tr, 1, t, 1);
          const uint32_t a_addr = p2r_smem_u32(smem + HALO_W_BYTES + s * halo_bytes);
#pragma unroll
          for (int tp = 0; tp < 3; ++tp) {
            const int off_rows = halo_v + tap.shift0 + tp * tap.shift_step;   // first row of this tap's 128-row view
            // A view that starts 25 rows into the TMA-written tile is NOT on the 1024-byte repeat of the 128-byte swizzle.
            // Measured on B200: the plain descriptor (matrix base offset 0) is exact, the one with base offset
            // (start >> 7) & 7 is wrong -- the swizzle XOR is a function of the absolute shared-memory address bits, for
            // the TMA write and for the MMA read alike, so any 128-byte-row offset just works.
            const uint32_t a_tap = a_addr + (uint32_t)off_rows * 128u;
            const uint32_t b_tap = p2r_smem_u32(smem + tp * 8192);
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
              const uint64_t adesc = make_desc(a_tap + k * 32, 16, 1024);
              const uint64_t bdesc = B_MN ? make_desc(b_tap + k * 16 * 128, GEMM_BLOCK_K * 128, 1024)
                                          : make_desc(b_tap + k * 32, 16, 1024);
              umma_bf16(tmem_acc, adesc, bdesc, idesc, (tp | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty + s);
          umma_commit(tmem_full + as);
          trace_mark(tr, 1, t, 2);
        }
      }
    } else if (TAP == 4) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(GEMM_BLOCK_M, 192, 1, 1);
        const uint32_t zero_addr = p2r_smem_u32(smem + S::STAGES * stage4_bytes);
        for (int i = 0; i < nkb; ++i) {
          const int s = i % S::STAGES;
          p2r_mbar_wait(full + s, (uint32_t)(i / S::STAGES) & 1u);
          tc_fence_after();
          const uint32_t a_addr = p2r_smem_u32(smem + s * stage4_bytes);
          const uint32_t b_addr = a_addr + DW4_A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // M block 1 (rows 64..127) = the shared zero block: leading byte offset = its distance from this stage's tile
            const uint64_t adesc = make_desc(a_addr + k * 16 * 128, zero_addr - a_addr, 1024);
            const uint64_t bdesc = make_desc(b_addr + k * 16 * 128, GEMM_BLOCK_K * 128, 1024);
            umma_bf16(tmem_base, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty + s);
        }
        umma_commit(tmem_full);
      }
    } else if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(GEMM_BLOCK_M, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int it = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int as = t % ACC;
        if (t >= ACC) {   // the epilogue must have drained this accumulator buffer
          p2r_mbar_wait(tmem_empty + as, (uint32_t)((t / ACC) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (uint32_t)(it / S::STAGES) & 1u;
          p2r_mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t a_addr = p2r_smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            const uint64_t adesc = A_MN ? make_desc(a_addr + k * 16 * 128, GEMM_BLOCK_K * 128, 1024)
                                        : make_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_desc(b_addr + k * 16 * 128, GEMM_BLOCK_K * 128, 1024)
                                        : make_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(tmem_acc, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty + s);        // stage reusable once these MMAs have read it
        }
        umma_commit(tmem_full + as);     // accumulator of tile t complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const bool vec_ok = ((size_t)ldc * sizeof(OutT)) % 16 == 0 && (reinterpret_cast<uintptr_t>(C) % 16 == 0);
    constexpr bool do_stats = STATS && !ATOMIC;   // (a template flag: the extra registers stay out of the plain kernels)
    if (TS) {
      // ---- TMA-store epilogue: TMEM -> registers -> bias / ReLU / bf16 -> swizzled 32 x 64 staging tile -> bulk store
      // PAIRED (TAP == 3, 8 epilogue warps): two warps share a TMEM lane quarter and its staging tile; warp `half` converts
      // the 32 columns [32 half, 32 half + 32) and sums rows [16 half, 16 half + 16) for the statistics -- the 4-warp
      // epilogue was the longest stage of the halo kernels (2 CTAs x 4 warps per SM, issue slots 45 % busy).
      constexpr bool PAIRED = TAP == 3;
      constexpr int EPI_THREADS = PAIRED ? 256 : 128;
      const int et = threadIdx.x - 64;
      const int half = PAIRED ? ((warp - 2) >> 2) : 0;
      const int ew = PAIRED ? (warp - 2) : q;                 // slot of this warp in the final reduction
      uint8_t* stg = staging + q * (32 * 128);
      float ts00 = 0.f, ts01 = 0.f, ts10 = 0.f, ts11 = 0.f;   // sum / sum of squares of channels 2 lane, 2 lane + 1
      long long* tr = (PAIRED && blockIdx.y == 1 && blockIdx.x == 0 && warp == 2 && lane == 0) ? g_tconv_trace : nullptr;
      // the bias of this CTA's column tile, once (every m-tile of the CTA has the same n0)
      const float* sb = sbias;
      for (int c = et; c < BLOCK_N; c += EPI_THREADS) sbias[c] = (bias != nullptr && n0 + c < N) ? __ldg(bias + n0 + c) : 0.f;
      if (PAIRED) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int t = 0; t < ntiles; ++t) {
        const int as = t % ACC;
        const int row0 = (tile0 + t) * GEMM_BLOCK_M + q * 32;
        const int ncols = min(BLOCK_N, N - n0);
        trace_mark(tr, 2, t, 0);
        trace_mark(tr, 2, t, 1);
        p2r_mbar_wait(tmem_full + as, (uint32_t)(t / ACC) & 1u);
        tc_fence_after();
        trace_mark(tr, 2, t, 2);
#pragma unroll 1
        for (int g0 = 0; g0 < ncols; g0 += 64) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (PAIRED && h != half) continue;
            const int c0 = g0 + 32 * h;
            float f[32];
            if (c0 < ncols) {   // (warp-uniform)
              uint32_t v[32];
              tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + c0), v);
              if (PAIRED || c0 + 32 >= ncols) {   // last read of this accumulator buffer (by this warp)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) p2r_mbar_arrive(tmem_empty + as);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = (nkb > 0 ? __uint_as_float(v[j]) : 0.f) + sb[c0 + j];
                if (relu) x = fmaxf(x, 0.f);
                f[j] = x;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = 0.f;
            }
            if (PAIRED) {   // the previous bulk store has read the staging tile and the partner is done summing it
              if (half == 0 && lane == 0) tma_store_wait_read();
              asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
              trace_mark(tr, 2, t, 3);
            } else if (h == 0) {   // the previous bulk store must have finished reading the staging tile
              if (lane == 0) tma_store_wait_read();
              __syncwarp();
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * p + 0], f[8 * p + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * p + 2], f[8 * p + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * p + 4], f[8 * p + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * p + 6], f[8 * p + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0);
              pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2);
              pk.w = *reinterpret_cast<uint32_t*>(&t3);
              const int piece = (4 * h + p) ^ (lane & 7);          // SWIZZLE_128B
              *reinterpret_cast<uint4*>(stg + lane * 128 + piece * 16) = pk;
            }
          }
          p2r_fence_proxy_async();
          __syncwarp();
          trace_mark(tr, 2, t, 4);
          if (PAIRED) asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");      // both halves of the tile are staged
          if (lane == 0 && half == 0) tma_store_2d(&tma_c, stg, n0 + g0, row0);   // rows >= M / columns >= N are clipped
          trace_mark(tr, 2, t, 5);
          if (do_stats && PAIRED) {
            // (all loads first: with a data-dependent trip count the compiler kept one shared-memory round trip per row,
            // 28 ns each in the in-kernel timeline)
            const int r_lo = 16 * half;
            const int rows_valid = min(32, M - row0);
            uint32_t wd[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int r = r_lo + i;
              wd[i] = r < rows_valid ? *reinterpret_cast<const uint32_t*>(stg + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2)) : 0u;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float lo = __uint_as_float(wd[i] << 16), hi = __uint_as_float(wd[i] & 0xffff0000u);
              ts00 += lo;
              ts01 = fmaf(lo, lo, ts01);
              ts10 += hi;
              ts11 = fmaf(hi, hi, ts11);
            }
          } else if (do_stats) {
            const int rows_valid = min(32, M - row0);
#pragma unroll 8
            for (int r = 0; r < rows_valid; ++r) {
              const uint32_t wd = *reinterpret_cast<const uint32_t*>(stg + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2));
              const float lo = __uint_as_float(wd << 16), hi = __uint_as_float(wd & 0xffff0000u);
              ts00 += lo;
              ts01 = fmaf(lo, lo, ts01);
              ts10 += hi;
              ts11 = fmaf(hi, hi, ts11);
            }
          }
          trace_mark(tr, 2, t, 6);
        }
      }
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
      tc_fence_before();
      if (do_stats) {
        float* red = reinterpret_cast<float*>(smem);   // the ring is idle: every MMA of this CTA has completed
        red[(ew * 4 + 0) * 32 + lane] = ts00;
        red[(ew * 4 + 1) * 32 + lane] = ts01;
        red[(ew * 4 + 2) * 32 + lane] = ts10;
        red[(ew * 4 + 3) * 32 + lane] = ts11;
        if (PAIRED) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 1, 128;" ::: "memory");
        if (half == 0) {
          const int k = q;
          float tot = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < (PAIRED ? 8 : 4); ++w8) tot += red[(w8 * 4 + k) * 32 + lane];
          const int copy = (int)((blockIdx.y * gridDim.x + blockIdx.x) % (unsigned)ex.stat_copies);
          atomicAdd(ex.stats + ((size_t)copy * 2 + (k & 1)) * 64 + 2 * lane + (k >> 1), (double)tot);
        }
      }
    } else {
    // statistics accumulators: st[h][r] belongs to channel 32 h + 16 r + (lane & 15); lanes < 16 hold the sum, lanes
    // >= 16 the sum of squares
    float st00 = 0.f, st01 = 0.f, st10 = 0.f, st11 = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int as = t % ACC;
      const int row = (tile0 + t) * GEMM_BLOCK_M + q * 32 + lane;
      p2r_mbar_wait(tmem_full + as, (uint32_t)(t / ACC) & 1u);
      tc_fence_after();
      const bool row_ok = row < M;
      OutT* crow = C + (size_t)row * ldc;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + c0), v);
        if (c0 + 32 >= BLOCK_N) {   // last read of this accumulator buffer: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) p2r_mbar_arrive(tmem_empty + as);
        }
        const int col0 = n0 + c0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = nkb > 0 ? __uint_as_float(v[j]) : 0.f;   // an n-tile without k-blocks is all zero
          if (!ATOMIC) {
            if (bias != nullptr && col0 + j < N) x += __ldg(bias + col0 + j);
            if (relu) x = fmaxf(x, 0.f);
            if (sizeof(OutT) == 2) x = __bfloat162float(__float2bfloat16_rn(x));   // the value as stored
          }
          f[j] = x;
        }
        if (row_ok && col0 < N) {  // (no `continue`: the next tcgen05.ld is warp-aligned)
          if (ATOMIC) {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) atomicAdd(reinterpret_cast<float*>(crow) + col0 + j, f[j]);
          } else if (vec_ok && col0 + 32 <= N) {
            if (sizeof(OutT) == 2) {
              uint4* dst = reinterpret_cast<uint4*>(crow + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * j + 0], f[8 * j + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                dst[j] = pk;
              }
            } else {
              float4* dst = reinterpret_cast<float4*>(crow + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) {
                if (sizeof(OutT) == 2) reinterpret_cast<__nv_bfloat16*>(crow)[col0 + j] = __float2bfloat16_rn(f[j]);
                else reinterpret_cast<float*>(crow)[col0 + j] = f[j];
              }
          }
        }
        if (do_stats) {   // warp-uniform: column sums of this 32 x 32 block (rows / columns outside the matrix count 0)
          // two rounds of 16 columns: c = {16 values, their 16 squares} -> one transposed 32-lane reduction leaves the
          // column sum on lane l < 16 and the sum of squares of the same column on lane l + 16
          const bool hi = (col0 >> 5) & 1;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            float c[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = (row_ok && col0 + 16 * r + j < N) ? f[16 * r + j] : 0.f;
              c[j] = x;
              c[16 + j] = x * x;
            }
            const float tot = warp_transpose_sum32(c, lane);
            if (hi) { if (r) st11 += tot; else st10 += tot; }
            else { if (r) st01 += tot; else st00 += tot; }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    if (do_stats) {
      // every MMA of this CTA has completed (tmem_full of the last tile), so the pipeline stages are free: combine
      // the four epilogue warps there, then 128 double atomics per CTA into this CTA's copy of the statistics
      float* red = reinterpret_cast<float*>(smem);   // [4 warps][4 values][32 lanes]
      red[(q * 4 + 0) * 32 + lane] = st00;
      red[(q * 4 + 1) * 32 + lane] = st01;
      red[(q * 4 + 2) * 32 + lane] = st10;
      red[(q * 4 + 3) * 32 + lane] = st11;
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      const int k = q;   // warp q finishes accumulator k = 2 h + r
      const float tot = red[(0 * 4 + k) * 32 + lane] + red[(1 * 4 + k) * 32 + lane] + red[(2 * 4 + k) * 32 + lane] +
                        red[(3 * 4 + k) * 32 + lane];
      const int copy = (int)((blockIdx.y * gridDim.x + blockIdx.x) % (unsigned)ex.stat_copies);
      atomicAdd(ex.stats + ((size_t)copy * 2 + (lane >> 4)) * 64 + k * 16 + (lane & 15), (double)tot);
    }
    tc_fence_before();
    }   // !TS
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, S::TMEM_COLS);
  }
}

// ================================================================================================ CTA-pair kernel
// gemm2_bf16_kernel: the same GEMM on a CTA PAIR (cluster of 2, tcgen05 cta_group::2): one MMA spans two SMs, M = 256
// rows (128 per CTA), N = BN columns.  Each CTA stages its own 128 x 64 A tile and only HALF of the B tile
// (BN/2 x 64), so a CTA reads (128 + BN/2) smem rows per k-block instead of (128 + BN): the single-CTA kernel above is
// bound by shared-memory bandwidth (TMA writes + MMA operand reads), not by the tensor pipe.
//   * both CTAs run a TMA producer; all transaction bytes land on the LEADER's full[] barrier (rank 0)
//   * only the leader issues tcgen05.mma; tcgen05.commit ... multicast::cluster releases the stage in BOTH CTAs and
//     finally signals tmem_full in both
//   * each CTA's epilogue warps drain their own TMEM half (128 lanes x BN columns) into their own 128 rows of C
// K-major A [M,K] and B [N,K], bf16 C, optional bias / ReLU / k-block lists / fused statistics; one tile per pair.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
  // executed by both CTAs of the pair; clearing the peer bit makes the bytes count on CTA 0's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          p2r_smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(p2r_smem_u32(leader_bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(p2r_smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {   // arrives on `bar` in both CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          p2r_smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive_on_leader(uint64_t* bar) {   // arrive on CTA 0's copy of `bar`
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(p2r_smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int BN>
struct PairSmem {
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BOX_ROWS = BN / 2;
  static constexpr int B_BYTES = B_BOX_ROWS * GEMM_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGING_BYTES = 4 * 32 * 128;                          // per epilogue warp: 32 rows x 64 bf16
  static constexpr int BIAS_BYTES = 2 * BN * 4;                               // bias of the tile, double-buffered
  static constexpr int MAX_STAGES = (232448 - 1024 - 256 - STAGING_BYTES - BIAS_BYTES) / STAGE_BYTES;   // 1 CTA per SM
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int ACC_COLS = 2 * BN;                                     // double-buffered accumulator
  static constexpr int TMEM_COLS = ACC_COLS <= 256 ? 256 : 512;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + BIAS_BYTES + 1024 + 256;
};

// Persistent: one pair per TPC (grid = 2 x 74), each pair walks tiles pair_id, pair_id + npairs, ... (n fastest, so
// the pairs working at the same time share the same rows of A in L2).  The smem ring runs across tile boundaries and
// the accumulator is double-buffered in TMEM, so the epilogue of tile t overlaps the loads and MMAs of tile t + 1.
template <int BN, bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_c, int M, int N, int K, const float* __restrict__ bias,
                  int relu, int tiles_n, int num_tiles, const GemmExtra ex) {
  using S = PairSmem<BN>;
  const bool dbg_nostore = (relu & 256) != 0, dbg_nomma = (relu & 512) != 0, dbg_nobias = (relu & 1024) != 0;
  // accumulate: C += A.B^T instead of C = A.B^T -- the epilogue's bulk-tensor stores become bulk-tensor REDUCE-ADD stores
  // (the bf16 addition happens in L2; same rounding as a separate elementwise add of two bf16 tensors).  The backward
  // pass uses it to add the graph convolution's input gradient onto the residual branch's gradient in place.
  const bool accumulate = (relu & 2) != 0;
  relu &= 1;
  if (dbg_nobias) bias = nullptr;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + S::STAGES * S::STAGE_BYTES;                       // 1024-byte aligned (stage sizes are)
  float* sbias = reinterpret_cast<float*>(staging + S::STAGING_BYTES);        // [2][BN]
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + S::STAGING_BYTES + S::BIAS_BYTES);
  uint64_t* empty = full + S::STAGES;
  uint64_t* tmem_full = empty + S::STAGES;    // [2]  signalled in both CTAs by the leader's commit
  uint64_t* tmem_empty = tmem_full + 2;       // [2]  used in the LEADER only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      p2r_mbar_init(full + s, 1);
      p2r_mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      p2r_mbar_init(tmem_full + a, 1);
      p2r_mbar_init(tmem_empty + a, 8);
    }
    p2r_fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();           // barriers of both CTAs initialised, both TMEM halves allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // per-tile geometry (shared by the three roles)
  auto tile_m0 = [&](int tile) { return (tile / tiles_n) * (2 * GEMM_BLOCK_M) + (int)rank * GEMM_BLOCK_M; };
  auto tile_effn = [&](int n0) {            // MMA width of this tile: the ragged last n-tile issues a narrower MMA
    const int rem = N - n0;
    return rem >= BN ? BN : ((rem + 15) & ~15);
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int it = 0;
      for (int tile = pair0; tile < num_tiles; tile += npairs) {
        const int n_tile = tile % tiles_n;
        const int n0 = n_tile * BN, m0 = tile_m0(tile);
        const int nb0 = n0 + (int)rank * (tile_effn(n0) / 2);             // this CTA's half of the B rows
        const int* kbl = ex.kb_list != nullptr ? ex.kb_list + (size_t)n_tile * ex.kb_stride + 1 : nullptr;
        const int nkb = kbl != nullptr ? __ldg(kbl - 1) : total_kb;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (uint32_t)(it / S::STAGES) & 1u;
          p2r_mbar_wait(empty + s, ph ^ 1u);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          if (rank == 0) p2r_mbar_expect_tx(full + s, 2 * S::STAGE_BYTES);   // the bytes of both CTAs
          const int k0 = (kbl != nullptr ? __ldg(kbl + i) : i) * GEMM_BLOCK_K;
          tma_load_2d_pair(a_dst, &tma_a, full + s, k0, m0);                 // box {64 k, 128 m}
          tma_load_2d_pair(b_dst, &tma_b, full + s, k0, nb0);                // box {64 k, BN/2 n}
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int it = 0, t = 0;
      for (int tile = pair0; tile < num_tiles; tile += npairs, ++t) {
        const int n_tile = tile % tiles_n;
        const int nkb = ex.kb_list != nullptr ? __ldg(ex.kb_list + (size_t)n_tile * ex.kb_stride) : total_kb;
        const uint32_t idesc = make_idesc(2 * GEMM_BLOCK_M, tile_effn(n_tile * BN), 0, 0);
        const int as = t & 1;
        if (t >= 2) {   // both CTAs' epilogues must have drained this accumulator buffer
          p2r_mbar_wait(tmem_empty + as, (uint32_t)((t >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (uint32_t)(it / S::STAGES) & 1u;
          p2r_mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t a_addr = p2r_smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t b_addr = a_addr + S::A_BYTES;
          if (!dbg_nomma) {
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
              umma_bf16_pair(tmem_acc, make_desc(a_addr + k * 32, 16, 1024), make_desc(b_addr + k * 32, 16, 1024), idesc,
                             (i | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(empty + s);      // stage reusable in both CTAs once these MMAs have read it
        }
        umma_commit_pair(tmem_full + as);   // accumulator of tile t complete in both CTAs
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    // TMEM -> registers (32 columns at a time) -> bias / ReLU / bf16 -> this warp's 32 x 64 staging tile in shared
    // memory (128-byte swizzle: conflict-free 16-byte writes) -> ONE bulk-tensor store per 64 columns.  Direct
    // st.global from this register layout (a lane = a row) writes 16-byte pieces 32 rows apart and was the bound of
    // the whole kernel; the TMA store writes full 128-byte lines and costs the warp one instruction.
    const int q = warp & 3;
    const int et = threadIdx.x - 64;            // 0..127 over the epilogue warps
    uint8_t* stg = staging + q * (32 * 128);
    float st00 = 0.f, st01 = 0.f, st10 = 0.f, st11 = 0.f;   // sum / sum of squares of channels 2 lane, 2 lane + 1
    int t = 0;
    for (int tile = pair0; tile < num_tiles; tile += npairs, ++t) {
      const int n_tile = tile % tiles_n;
      const int n0 = n_tile * BN;
      const int nkb = ex.kb_list != nullptr ? __ldg(ex.kb_list + (size_t)n_tile * ex.kb_stride) : total_kb;
      const int as = t & 1;
      const int row0 = tile_m0(tile) + q * 32;
      const int ncols = min(BN, N - n0);
      float* sb = sbias + as * BN;
      for (int c = et; c < BN; c += 128) sb[c] = (bias != nullptr && n0 + c < N) ? __ldg(bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");   // bias of this tile visible to the four epilogue warps
      p2r_mbar_wait(tmem_full + as, (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int g0 = 0; g0 < ncols; g0 += 64) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = g0 + 32 * h;
          float f[32];
          if (c0 < ncols) {   // (warp-uniform)
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), v);
            if (c0 + 32 >= ncols) {   // last read of this accumulator buffer: hand it back to the leader's MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_on_leader(tmem_empty + as);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = (nkb > 0 ? __uint_as_float(v[j]) : 0.f) + sb[c0 + j];
              if (relu) x = fmaxf(x, 0.f);
              f[j] = __bfloat162float(__float2bfloat16_rn(x));   // the value as stored
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;
          }
          if (h == 0) {   // the previous bulk store must have finished reading the staging tile
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
          }
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            uint4 pk;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * p + 0], f[8 * p + 1]);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * p + 2], f[8 * p + 3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * p + 4], f[8 * p + 5]);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * p + 6], f[8 * p + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&t0);
            pk.y = *reinterpret_cast<uint32_t*>(&t1);
            pk.z = *reinterpret_cast<uint32_t*>(&t2);
            pk.w = *reinterpret_cast<uint32_t*>(&t3);
            const int piece = (4 * h + p) ^ (lane & 7);          // SWIZZLE_128B: 16-byte chunk index ^ (row & 7)
            *reinterpret_cast<uint4*>(stg + lane * 128 + piece * 16) = pk;
          }
        }
        p2r_fence_proxy_async();      // generic-proxy writes of the staging tile -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0 && !dbg_nostore) {                                             // rows >= M / columns >= N are clipped
          if (accumulate) tma_reduce_add_2d(&tma_c, stg, n0 + g0, row0);
          else tma_store_2d(&tma_c, stg, n0 + g0, row0);
        }
        if (STATS) {
          // column sums straight from the staged bf16 tile: lane l owns columns 2l, 2l+1 of this 64-column group
          // (= channels 2l, 2l+1: n0 + g0 is a multiple of 64); one conflict-free 4-byte read per row
          // (eight loads, then their sums: with the row count as the trip count the compiler kept one shared-memory round
          // trip per row -- 28 ns each in the temporal-conv timeline)
          const int rows_valid = min(32, M - row0);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint32_t wd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = c8 * 8 + i;
              wd[i] = r < rows_valid ? *reinterpret_cast<const uint32_t*>(stg + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + ((lane & 3) << 2)) : 0u;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float lo = __uint_as_float(wd[i] << 16), hi = __uint_as_float(wd[i] & 0xffff0000u);
              st00 += lo;
              st01 = fmaf(lo, lo, st01);
              st10 += hi;
              st11 = fmaf(hi, hi, st11);
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
    tc_fence_before();
    if (STATS) {
      float* red = reinterpret_cast<float*>(smem);   // the ring is idle: every MMA of this pair has completed
      red[(q * 4 + 0) * 32 + lane] = st00;
      red[(q * 4 + 1) * 32 + lane] = st01;
      red[(q * 4 + 2) * 32 + lane] = st10;
      red[(q * 4 + 3) * 32 + lane] = st11;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int k = q;
      const float tot = red[(0 * 4 + k) * 32 + lane] + red[(1 * 4 + k) * 32 + lane] + red[(2 * 4 + k) * 32 + lane] +
                        red[(3 * 4 + k) * 32 + lane];
      const int copy = (int)(blockIdx.x % (unsigned)ex.stat_copies);
      // accumulator k: 0 = sum of channel 2l, 1 = its sum of squares, 2 / 3 = the same for channel 2l + 1
      atomicAdd(ex.stats + ((size_t)copy * 2 + (k & 1)) * 64 + 2 * lane + (k >> 1), (double)tot);
    }
  }
  tc_fence_before();
  cluster_sync_all();            // neither CTA leaves (or frees TMEM) while the other may still use the pair's state
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, S::TMEM_COLS);
  }
}

// gemm2_dw_kernel: weight gradient of the graph convolution on CTA pairs.
//   dW[N1, N2] (fp32, zero-filled by the caller) += dz^T . x,   dz [R, N1], x [R, N2] bf16 row-major (both operands are
//   "MN-major": the reduction runs over the ROWS R = B*T frames), 256 x 256 output tiles, structurally-zero tiles are
//   not in `tile_list`, the reduction is split `splits` ways; a work unit u = (split u / num_tiles, tile u % num_tiles)
//   so the pairs working at the same time stream the same rows of dz and x (shared through L2).  The epilogue stages
//   32 x 32 fp32 tiles in shared memory and adds them into dW with bulk-tensor REDUCE stores (cp.reduce.async.bulk
//   .add): split-K without per-element atomics.
template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_dw_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ CUtensorMap tma_c, int N1, int N2, int R, const int* __restrict__ tile_list,
                int num_tiles, int splits, int kps) {
  using S = PairSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + S::STAGES * S::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + S::STAGING_BYTES + S::BIAS_BYTES);
  uint64_t* empty = full + S::STAGES;
  uint64_t* tmem_full = empty + S::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total_kb = (R + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int num_units = num_tiles * splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      p2r_mbar_init(full + s, 1);
      p2r_mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      p2r_mbar_init(tmem_full + a, 1);
      p2r_mbar_init(tmem_empty + a, 8);
    }
    p2r_fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_c);
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, S::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto unit_effn = [&](int n0) {            // MMA width: whole 64-column blocks per CTA (a multiple of 128 in total)
    const int rem = N2 - n0;
    return rem >= BN ? BN : ((rem + 127) & ~127);
  };
  auto unit_kb0 = [&](int u) { return (u / num_tiles) * kps; };
  auto unit_nkb = [&](int u) { return min(total_kb, unit_kb0(u) + kps) - unit_kb0(u); };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int it = 0;
      for (int u = pair0; u < num_units; u += npairs) {
        const int tile = u % num_tiles;
        const int m0 = __ldg(tile_list + 2 * tile) * (2 * GEMM_BLOCK_M) + (int)rank * GEMM_BLOCK_M;
        const int n0 = __ldg(tile_list + 2 * tile + 1) * BN;
        const int nb0 = n0 + (int)rank * (unit_effn(n0) / 2);
        const int kb0 = unit_kb0(u), nkb = unit_nkb(u);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (uint32_t)(it / S::STAGES) & 1u;
          p2r_mbar_wait(empty + s, ph ^ 1u);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          if (rank == 0) p2r_mbar_expect_tx(full + s, 2 * S::STAGE_BYTES);
          const int k0 = (kb0 + i) * GEMM_BLOCK_K;
#pragma unroll
          for (int h = 0; h < GEMM_BLOCK_M / 64; ++h)                      // boxes {64 m, 64 k}
            tma_load_2d_pair(a_dst + h * (GEMM_BLOCK_K * 128), &tma_a, full + s, m0 + h * 64, k0);
#pragma unroll
          for (int h = 0; h < BN / 128; ++h)                               // boxes {64 n, 64 k}
            tma_load_2d_pair(b_dst + h * (GEMM_BLOCK_K * 128), &tma_b, full + s, nb0 + h * 64, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int it = 0, t = 0;
      for (int u = pair0; u < num_units; u += npairs, ++t) {
        const int tile = u % num_tiles;
        const int n0 = __ldg(tile_list + 2 * tile + 1) * BN;
        const uint32_t idesc = make_idesc(2 * GEMM_BLOCK_M, unit_effn(n0), 1, 1);
        const int nkb = unit_nkb(u);
        const int as = t & 1;
        if (t >= 2) {
          p2r_mbar_wait(tmem_empty + as, (uint32_t)((t >> 1) - 1) & 1u);
          tc_fence_after();
        }
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BN);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % S::STAGES;
          const uint32_t ph = (uint32_t)(it / S::STAGES) & 1u;
          p2r_mbar_wait(full + s, ph);
          tc_fence_after();
          const uint32_t a_addr = p2r_smem_u32(smem + s * S::STAGE_BYTES);
          const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
            umma_bf16_pair(tmem_acc, make_desc(a_addr + k * 16 * 128, GEMM_BLOCK_K * 128, 1024),
                           make_desc(b_addr + k * 16 * 128, GEMM_BLOCK_K * 128, 1024), idesc, (i | k) != 0 ? 1u : 0u);
          umma_commit_pair(empty + s);
        }
        umma_commit_pair(tmem_full + as);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs): fp32 tiles, bulk-tensor reduce-add =====================
    const int q = warp & 3;
    uint8_t* stg = staging + q * (32 * 128);
    int t = 0;
    for (int u = pair0; u < num_units; u += npairs, ++t) {
      const int tile = u % num_tiles;
      const int row0 = __ldg(tile_list + 2 * tile) * (2 * GEMM_BLOCK_M) + (int)rank * GEMM_BLOCK_M + q * 32;
      const int n0 = __ldg(tile_list + 2 * tile + 1) * BN;
      const int ncols = min(BN, N2 - n0);
      const int nkb = unit_nkb(u);
      const int as = t & 1;
      p2r_mbar_wait(tmem_full + as, (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c0), v);
        if (c0 + 32 >= ncols) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_on_leader(tmem_empty + as);
        }
        if (lane == 0) tma_store_wait_read();     // the previous reduce-store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          uint4 pk;
          pk.x = nkb > 0 ? v[4 * p + 0] : 0u;
          pk.y = nkb > 0 ? v[4 * p + 1] : 0u;
          pk.z = nkb > 0 ? v[4 * p + 2] : 0u;
          pk.w = nkb > 0 ? v[4 * p + 3] : 0u;
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((p ^ (lane & 7)) << 4)) = pk;   // SWIZZLE_128B
        }
        p2r_fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row0 < N1) tma_reduce_add_2d(&tma_c, stg, n0 + c0, row0);   // rows / columns outside dW are clipped
      }
    }
    if (lane == 0) tma_store_wait_all();
    __syncwarp();
    tc_fence_before();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, S::TMEM_COLS);
  }
}


extern "C" int p2r_debug_tconv_trace(long long* device_buffer) {
  return cudaMemcpyToSymbol(g_tconv_trace, &device_buffer, sizeof(device_buffer)) == cudaSuccess ? 0 : -1;
}

// P2R_TCONV_HALO: 1 (default) the halo-tile forward / input-gradient kernels and the six-stage weight gradient, 0 the
// three-view kernels of round 1
static int tconv_halo_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("P2R_TCONV_HALO");
    v = e == nullptr ? 1 : atoi(e);
  }
  return v;
}

static bool ts_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("P2R_TMA_STORE");
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int TAP, int NSTAGE>
static int launch_with_maps(const CUtensorMap& ma, const CUtensorMap& mb, void* C, int ldc, int c_dtype, int M, int N,
                            int K, const float* bias, int relu, int splits, const TapArgs& tap, cudaStream_t st,
                            const GemmExtra& ex = GemmExtra{nullptr, 0, nullptr, nullptr, 1}) {
  const int total_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  int kps = total_kb;
  if (splits > 1) {
    kps = (total_kb + splits - 1) / splits;
    splits = (total_kb + kps - 1) / kps;
  } else splits = 1;
  // TMA-store epilogue for bf16 outputs of the 64- and 256-wide tilings (the memory-bound point-MLP / temporal-conv
  // GEMMs; with 128 / 160 columns the extra staging memory would cost the second co-resident CTA)
  constexpr bool TS_OK = (BLOCK_N == 64 || BLOCK_N == 256) && TAP != 2 && TAP != 4;
  const bool ts = TS_OK && c_dtype == 1 && splits == 1 && ex.tile_mask == nullptr && ldc % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(C) % 16) == 0 && ts_enabled();
  if (TAP == 3 && !ts) {
    p2r_set_last_error("p2r_tconv_bf16: the halo kernels need the TMA-store epilogue (bf16 output, 16-byte aligned)", -1);
    return -1;
  }
  CUtensorMap mc = ma;
  if (ts && make_map(&mc, C, N, M, ldc, 32)) return -1;
  // several m-tiles per CTA when there are many more tiles than CTA slots: barrier / TMEM set-up is amortised and the
  // next tile's TMA loads and MMAs run under the current tile's epilogue (double-buffered accumulator)
  const long long tiles_m = p2r_ceil_div(M, GEMM_BLOCK_M), tiles_n = p2r_ceil_div(N, BLOCK_N);
  int tpc = 1;
  if (splits == 1 && GemmSmem<BLOCK_N, NSTAGE>::ACC_STAGES == 2 && ex.tile_mask == nullptr) {
    const long long slots = (long long)P2R_SM_COUNT * 8;
    tpc = (int)((tiles_m * tiles_n) / slots);
    if (tpc < 1) tpc = 1;
    if (tpc > 8) tpc = 8;
  }
  if (TAP == 3) {   // persistent: one wave of 2 CTAs per SM walks all tiles (set-up / weights / tail once per CTA)
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("P2R_TCONV_TPC");
      forced = e == nullptr ? 0 : atoi(e);
    }
    tpc = forced > 0 ? forced : (int)p2r_ceil_div(tiles_m * tiles_n, (long long)P2R_SM_COUNT * 2);
    if (tpc < 1) tpc = 1;
  }
  dim3 grid((unsigned)tiles_n, (unsigned)p2r_ceil_div(tiles_m, tpc), splits);
#define GEMM_GO(OutT, ATOMIC, STATS, TS)                                                                        \
  do {                                                                                                          \
    using SS = GemmSmem<BLOCK_N, NSTAGE, TS>;                                                                   \
    auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, OutT, ATOMIC, TAP, NSTAGE, STATS, TS>;                    \
    const int v_rows = tap.shift_step < 0 ? -tap.shift_step : tap.shift_step;                                   \
    const int smem_bytes = TAP == 3 ? HALO_W_BYTES + HALO_STAGES * halo_tile_bytes(v_rows) + SS::STAGING_BYTES + \
                                          SS::BIAS_BYTES + 1024 + 256                                           \
                         : (TAP == 4 ? SS::STAGES * DW4_STAGE_BYTES + DW4_A_BYTES + 1024 + 256                  \
                                     : SS::TOTAL);                                                              \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);                        \
    kern<<<grid, TAP == 3 ? GEMM_THREADS + 128 : GEMM_THREADS, smem_bytes, st>>>(ma, mb, mc, (OutT*)C, ldc, M, N, K, bias, relu, kps, tap, tpc, ex); \
  } while (0)
  if (ex.stats != nullptr) {
    if (A_MN || B_MN || TAP == 2 || TAP == 4 || splits > 1 || c_dtype != 1) {
      p2r_set_last_error("p2r_gemm_bf16_ex: fused statistics need K-major operands, bf16 output, no split-K", -1);
      return -1;
    }
    if (!A_MN && !B_MN && TAP != 2 && TAP != 4) {
      if (TS_OK && ts) GEMM_GO(__nv_bfloat16, false, true, TS_OK);
      else GEMM_GO(__nv_bfloat16, false, true, false);
    }
  } else if (splits > 1) GEMM_GO(float, true, false, false);
  else if (c_dtype == 1) {
    if (TS_OK && ts) GEMM_GO(__nv_bfloat16, false, false, TS_OK);
    else GEMM_GO(__nv_bfloat16, false, false, false);
  } else GEMM_GO(float, false, false, false);
#undef GEMM_GO
  P2R_RETURN_LAUNCH("p2r_gemm_bf16");
}

template <int BLOCK_N, bool A_MN, bool B_MN>
static int launch_gemm(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int c_dtype, int M, int N,
                       int K, const float* bias, int relu, int splits, cudaStream_t st, const GemmExtra& ex) {
  CUtensorMap ma, mb;
  // K-major: rows = M (or N), inner = K.   MN-major: rows = K, inner = M (or N).
  if (make_map(&ma, A, A_MN ? M : K, A_MN ? K : M, lda, A_MN ? GEMM_BLOCK_K : GEMM_BLOCK_M)) return -1;
  if (make_map(&mb, B, B_MN ? N : K, B_MN ? K : N, ldb, B_MN ? GEMM_BLOCK_K : BLOCK_N)) return -1;
  const TapArgs none = {1, 1, 1, 0, 0};
  const int total_kb = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  // short reductions (the 64-channel point MLPs): fewer stages -> less smem -> more co-resident CTAs per SM to hide
  // the per-tile latency chain (TMEM alloc, TMA, MMA, epilogue)
  if (BLOCK_N == 64 && splits <= 1 && total_kb <= 1)
    return launch_with_maps<BLOCK_N == 64 ? 64 : BLOCK_N, A_MN, B_MN, 0, (BLOCK_N == 64 ? 1 : 0)>(ma, mb, C, ldc, c_dtype, M, N, K, bias, relu, splits, none, st, ex);
  if (BLOCK_N == 64 && splits <= 1 && total_kb <= 3)
    return launch_with_maps<BLOCK_N == 64 ? 64 : BLOCK_N, A_MN, B_MN, 0, (BLOCK_N == 64 ? 2 : 0)>(ma, mb, C, ldc, c_dtype, M, N, K, bias, relu, splits, none, st, ex);
  // tall split-K weight gradients of the 64-channel layers (dW[64,64] over ~800 k rows, one CTA per SM): the reduction is
  // pure streaming, and with 4 stages of 16 KB real data per SM it ran at ~27 GB/s per SM -- 8 stages keep twice as much
  // in flight
  if constexpr (BLOCK_N == 64 && A_MN && B_MN) {
    if (splits > 1 && total_kb >= 8 * splits)
      return launch_with_maps<64, true, true, 0, 8>(ma, mb, C, ldc, c_dtype, M, N, K, bias, relu, splits, none, st, ex);
  }
  return launch_with_maps<BLOCK_N, A_MN, B_MN, 0, 0>(ma, mb, C, ldc, c_dtype, M, N, K, bias, relu, splits, none, st, ex);
}

// (KT x 1) temporal convolution, zero padding (KT-1)/2, as an IMPLICIT GEMM on the 3-D activation tensor
// (ref: nn.Conv2d(C, C, (3,1), padding (1,0)) in st_gcn_block.tcn, stgcn_layers.py:405-411).  No unfold buffer.
//   mode 0 (forward) : act = x  [B, rows, Ci] bf16, w = W2 [Co, KT*Ci] (column dt*Ci+ci)  -> out y  [B*rows, Co] bf16 (+bias)
//   mode 1 (d input) : act = dy [B, rows, Co] bf16, w = Wt [KT*Co, Ci] (row dt*Co+co)     -> out dx [B*rows, Ci] bf16
//   mode 2 (d weight): act = x  [B, rows, Ci] bf16, other = dy [B*rows, Co] bf16          -> out dW2 [Co, KT*Ci] fp32
//                      (zero-filled by the caller when splits > 1)
// rows = T*V rows per sample, a tap shifts by V rows; requires rows % 128 == 0, Ci % 64 == 0, Co % 64 == 0, Co,Ci <= 64*...
extern "C" int p2r_tconv_bf16(int mode, const void* act, const void* w, const void* other, void* out, int B, int rows,
                              int Ci, int Co, int KT, int V, const float* bias, int splits, double* stats,
                              int stat_copies, void* stream) {
  P2R_CHECK_ARG(stats == nullptr || (mode == 0 && stat_copies >= 1), "p2r_tconv_bf16 (fused output statistics: forward only)");
  P2R_CHECK_ARG(mode >= 0 && mode <= 2 && B > 0 && rows > 0 && (rows % 128) == 0 && (KT & 1) && KT >= 1,
                "p2r_tconv_bf16");
  P2R_CHECK_ARG(Ci == 64 && Co == 64, "p2r_tconv_bf16 (built for the 64 -> 64 channel temporal conv of the hot path)");
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = (KT - 1) / 2;
  const int M = B * rows;
  CUtensorMap ma, mb;
  // one halo tile per 128 output rows instead of three shifted views (see TAP == 3 above): KT = 3, box <= 256 rows
  const bool halo = KT == 3 && V >= 1 && GEMM_BLOCK_M + 2 * V <= 256 && tconv_halo_mode() != 0 && ts_enabled() &&
                    (reinterpret_cast<uintptr_t>(out) % 16) == 0;
  if (mode == 0) {
    if (make_map3(&ma, act, Ci, rows, B, halo ? GEMM_BLOCK_M + 2 * V : GEMM_BLOCK_M)) return -1;
    if (make_map(&mb, w, (long long)KT * Ci, Co, (long long)KT * Ci, 64)) return -1;
    const TapArgs tap = {Ci / 64, rows, Ci, -pad * V, V};
    const GemmExtra ex = {nullptr, 0, nullptr, stats, stat_copies};
    if (halo) return launch_with_maps<64, false, false, 3, 4>(ma, mb, out, Co, 1, M, Co, KT * Ci, bias, 0, 1, tap, st, ex);
    return launch_with_maps<64, false, false, 1, 2>(ma, mb, out, Co, 1, M, Co, KT * Ci, bias, 0, 1, tap, st, ex);
  }
  if (mode == 1) {
    if (make_map3(&ma, act, Co, rows, B, halo ? GEMM_BLOCK_M + 2 * V : GEMM_BLOCK_M)) return -1;
    if (make_map(&mb, w, Ci, (long long)KT * Co, Ci, GEMM_BLOCK_K)) return -1;
    const TapArgs tap = {Co / 64, rows, Co, pad * V, -V};
    if (halo) return launch_with_maps<64, false, true, 3, 4>(ma, mb, out, Ci, 1, M, Ci, KT * Co, nullptr, 0, 1, tap, st);
    return launch_with_maps<64, false, true, 1, 2>(ma, mb, out, Ci, 1, M, Ci, KT * Co, nullptr, 0, 1, tap, st);
  }
  // mode 2: dW2[Co, KT*Ci] = dy^T . shifted(x)
  const bool deep_dw = KT == 3 && tconv_halo_mode() != 0;
  if (make_map(&ma, other, Co, M, Co, GEMM_BLOCK_K)) return -1;          // dy as MN-major A: inner = Co, rows = m
  if (make_map3(&mb, act, Ci, rows, B, GEMM_BLOCK_K)) return -1;         // x as MN-major B, 3-D
  const TapArgs tap = {1, rows, Ci, -pad * V, V};
  if (deep_dw) return launch_with_maps<192, true, true, 4, 6>(ma, mb, out, KT * Ci, 0, Co, KT * Ci, M, nullptr, 0, splits, tap, st);
  // KT = 3: ONE 192-column tile holds all three taps, so dy is streamed once (not once per tap) and a split is a single
  // CTA: 32 KB instead of 48 KB through L2 -> SM per 64 reduction rows
  if (KT == 3) return launch_with_maps<192, true, true, 2, 4>(ma, mb, out, KT * Ci, 0, Co, KT * Ci, M, nullptr, 0, splits, tap, st);
  return launch_with_maps<64, true, true, 2, 0>(ma, mb, out, KT * Ci, 0, Co, KT * Ci, M, nullptr, 0, splits, tap, st);
}

// C[M,N] (+)= op(A) . op(B)^T with bf16 operands (see file header).
//   a_mn = 0: A is [M,K] row-major (lda);  a_mn = 1: A is [K,M] row-major (lda)
//   b_mn = 0: B is [N,K] row-major (ldb);  b_mn = 1: B is [K,N] row-major (ldb)
//   c_dtype 0 = fp32, 1 = bf16;  splits > 1: split-K, fp32 atomics into a zero-filled C (no bias / ReLU)
//   block_n in {64, 128, 160, 256} (160 only with b_mn = 0); 0 = choose.
//   kb_list / kb_stride, tile_mask, stats / stat_copies: see GemmExtra (all optional; need splits <= 1 and an explicit
//   block_n, because the tables are indexed by this launch's tiling; stats need N % 64 == 0 channels-last columns).
extern "C" int p2r_gemm_bf16_ex(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn,
                                void* C, int ldc, int c_dtype, const float* bias, int relu, int splits, int block_n,
                                const int* kb_list, int kb_stride, const unsigned char* tile_mask, double* stats,
                                int stat_copies, void* stream) {
  P2R_CHECK_ARG(M > 0 && N > 0 && K > 0, "p2r_gemm_bf16");
  const bool extras = kb_list != nullptr || tile_mask != nullptr || stats != nullptr;
  P2R_CHECK_ARG(!extras || (splits <= 1 && block_n != 0), "p2r_gemm_bf16_ex (extras need splits <= 1 and an explicit block_n)");
  P2R_CHECK_ARG(kb_list == nullptr || kb_stride >= 1, "p2r_gemm_bf16_ex kb_stride");
  P2R_CHECK_ARG(stats == nullptr || (stat_copies >= 1 && N % 64 == 0), "p2r_gemm_bf16_ex stats");
  const GemmExtra ex = {kb_list, kb_stride, tile_mask, stats, stat_copies < 1 ? 1 : stat_copies};
  P2R_CHECK_ARG(!(splits > 1 && (c_dtype != 0 || bias || relu)), "p2r_gemm_bf16 (split-K needs fp32 C, no epilogue)");
  P2R_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "p2r_gemm_bf16 (row pitches must be multiples of 8 bf16 = 16 bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  if (block_n == 0) {
    if (!b_mn && N % 160 == 0 && N >= 640) block_n = 160;
    else if (N <= 64) block_n = 64;
    else if (N <= 128 || (N % 256 != 0 && N % 128 == 0) || N < 512) block_n = 128;
    else block_n = 256;
  }
  P2R_CHECK_ARG(block_n == 64 || block_n == 128 || block_n == 256 || (block_n == 160 && !b_mn), "p2r_gemm_bf16 block_n");
#define GEMM_DISPATCH(BN)                                                                                          \
  do {                                                                                                             \
    if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(A, lda, B, ldb, C, ldc, c_dtype, M, N, K, bias, relu, splits, st, ex); \
    if (!a_mn && b_mn) return launch_gemm<BN, false, true>(A, lda, B, ldb, C, ldc, c_dtype, M, N, K, bias, relu, splits, st, ex);  \
    if (a_mn && !b_mn) return launch_gemm<BN, true, false>(A, lda, B, ldb, C, ldc, c_dtype, M, N, K, bias, relu, splits, st, ex);  \
    return launch_gemm<BN, true, true>(A, lda, B, ldb, C, ldc, c_dtype, M, N, K, bias, relu, splits, st, ex);      \
  } while (0)
  if (block_n == 64) GEMM_DISPATCH(64);
  if (block_n == 128) GEMM_DISPATCH(128);
  if (block_n == 256) GEMM_DISPATCH(256);
  // 160: K-major B only
  if (!a_mn) return launch_gemm<160, false, false>(A, lda, B, ldb, C, ldc, c_dtype, M, N, K, bias, relu, splits, st, ex);
  return launch_gemm<160, true, false>(A, lda, B, ldb, C, ldc, c_dtype, M, N, K, bias, relu, splits, st, ex);
#undef GEMM_DISPATCH
}

extern "C" int p2r_gemm_bf16(int M, int N, int K, const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn,
                             void* C, int ldc, int c_dtype, const float* bias, int relu, int splits, int block_n,
                             void* stream) {
  return p2r_gemm_bf16_ex(M, N, K, A, lda, a_mn, B, ldb, b_mn, C, ldc, c_dtype, bias, relu, splits, block_n, nullptr, 0,
                          nullptr, nullptr, 1, stream);
}

template <int BN>
static int launch_pair(const void* A, int lda, const void* B, int ldb, void* C, int ldc, int M, int N, int K,
                       const float* bias, int relu, const GemmExtra& ex, cudaStream_t st) {
  using S = PairSmem<BN>;
  CUtensorMap ma, mb;
  CUtensorMap mc;
  if (make_map(&ma, A, K, M, lda, GEMM_BLOCK_M)) return -1;
  if (make_map(&mb, B, K, N, ldb, BN / 2)) return -1;
  if (make_map(&mc, C, N, M, ldc, 32)) return -1;                        // store box {64 columns, 32 rows}
  const int tiles_n = p2r_ceil_div(N, BN), pairs_m = p2r_ceil_div(M, 2 * GEMM_BLOCK_M);
  const int num_tiles = tiles_n * pairs_m;
  const unsigned grid = 2u * (unsigned)(num_tiles < P2R_SM_COUNT / 2 ? num_tiles : P2R_SM_COUNT / 2);   // one pair per TPC
  if (ex.stats != nullptr) {
    auto kern = gemm2_bf16_kernel<BN, true>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    kern<<<grid, GEMM_THREADS, S::TOTAL, st>>>(ma, mb, mc, M, N, K, bias, relu, tiles_n, num_tiles, ex);
  } else {
    auto kern = gemm2_bf16_kernel<BN, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    kern<<<grid, GEMM_THREADS, S::TOTAL, st>>>(ma, mb, mc, M, N, K, bias, relu, tiles_n, num_tiles, ex);
  }
  P2R_RETURN_LAUNCH("p2r_gemm_bf16_pair");
}


// CTA-pair (cta_group::2) variant for the large K-major GEMMs of the graph convolution: C[M,N] bf16 = A[M,K] . B[N,K]^T
// (+bias)(ReLU), 256 x block_n tile per pair, block_n in {128, 256}; kb_list / stats as in p2r_gemm_bf16_ex
// (kb_list indexed by n-tiles of width block_n).
extern "C" int p2r_gemm_bf16_pair(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                                  const float* bias, int relu, int block_n, const int* kb_list, int kb_stride,
                                  double* stats, int stat_copies, void* stream) {
  P2R_CHECK_ARG(M > 0 && N > 0 && K > 0, "p2r_gemm_bf16_pair");
  P2R_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0,
                "p2r_gemm_bf16_pair (row pitches must be multiples of 8 bf16 = 16 bytes)");
  P2R_CHECK_ARG(block_n == 128 || block_n == 256, "p2r_gemm_bf16_pair block_n (a multiple of the 64-column store box)");
  P2R_CHECK_ARG(kb_list == nullptr || kb_stride >= 1, "p2r_gemm_bf16_pair kb_stride");
  P2R_CHECK_ARG(stats == nullptr || (stat_copies >= 1 && N % 64 == 0), "p2r_gemm_bf16_pair stats");
  const GemmExtra ex = {kb_list, kb_stride, nullptr, stats, stat_copies < 1 ? 1 : stat_copies};
  cudaStream_t st = (cudaStream_t)stream;
  if (block_n == 128) return launch_pair<128>(A, lda, B, ldb, C, ldc, M, N, K, bias, relu, ex, st);
  return launch_pair<256>(A, lda, B, ldb, C, ldc, M, N, K, bias, relu, ex, st);
}

// 2-D fp32 tensor map over C [rows, cols] (ld floats between rows): box {32 columns = 128 bytes, 32 rows}, 128B swizzle.
static int make_map_f32(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { p2r_set_last_error("p2r_gemm_bf16_pair_dw: cuTensorMapEncodeTiled entry point unavailable", -1); return -1; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { p2r_set_last_error("p2r_gemm_bf16_pair_dw: cuTensorMapEncodeTiled (fp32) failed", -1); return -1; }
  return 0;
}

// Weight gradient on CTA pairs: dW[N1, N2] fp32 (zero-filled by the caller) += dz^T . x with dz [R, N1], x [R, N2] bf16
// row-major.  tile_list: int pairs (row tile, column tile) of the 256 x 256 output tiles to compute (the structurally
// non-zero ones), num_tiles of them; splits: how many ways the reduction over R is split (reduce-add stores combine them).
extern "C" int p2r_gemm_bf16_pair_dw(int R, int N1, int N2, const void* dz, int ldz, const void* x, int ldx, float* dW,
                                     int ldw, const int* tile_list, int num_tiles, int splits, void* stream) {
  P2R_CHECK_ARG(R > 0 && N1 > 0 && N2 > 0 && num_tiles > 0 && tile_list != nullptr && splits >= 1, "p2r_gemm_bf16_pair_dw");
  P2R_CHECK_ARG(ldz % 8 == 0 && ldx % 8 == 0 && ldw % 4 == 0, "p2r_gemm_bf16_pair_dw (row pitches must be multiples of 16 bytes)");
  constexpr int BN = 256;
  using S = PairSmem<BN>;
  CUtensorMap ma, mb, mc;
  if (make_map(&ma, dz, N1, R, ldz, GEMM_BLOCK_K)) return -1;     // MN-major operands: inner = output index, rows = reduction
  if (make_map(&mb, x, N2, R, ldx, GEMM_BLOCK_K)) return -1;
  if (make_map_f32(&mc, dW, N2, N1, ldw)) return -1;
  const int total_kb = (R + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  int kps = (total_kb + splits - 1) / splits;
  splits = (total_kb + kps - 1) / kps;
  const long long units = (long long)num_tiles * splits;
  // The weight gradient runs on a second stream beside the HBM-bound BatchNorm chain; a persistent kernel that took every
  // SM (226 KB of shared memory, all of TMEM) would shut that chain out, so it is given a SHARE of the CTA pairs.
  static int dw_pairs = 0;
  if (dw_pairs == 0) {
    const char* e = getenv("P2R_DW_PAIRS");
    dw_pairs = e ? atoi(e) : P2R_SM_COUNT / 2;
    if (dw_pairs < 1 || dw_pairs > P2R_SM_COUNT / 2) dw_pairs = P2R_SM_COUNT / 2;
  }
  const unsigned grid = 2u * (unsigned)(units < dw_pairs ? units : dw_pairs);
  auto kern = gemm2_dw_kernel<BN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
  kern<<<grid, GEMM_THREADS, S::TOTAL, (cudaStream_t)stream>>>(ma, mb, mc, N1, N2, R, tile_list, num_tiles, splits, kps);
  P2R_RETURN_LAUNCH("p2r_gemm_bf16_pair_dw");
}
