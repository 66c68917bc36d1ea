// p2r_common.cuh -- shared device helpers for the sm_100a kernels of pose2room_b200.
#pragma once
#ifdef P2R_HOST_EMULATION
#include "cuda_emu.h"   // tests/csrc: host emulator for the SIMT-only kernels (tests/test_kernels_emulated.py), never shipped
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#endif
#include <stdint.h>

#ifndef P2R_SM_COUNT        // (the host-emulation tests shrink it so that persistent kernels wrap their rings on tiny inputs)
#define P2R_SM_COUNT 148
#endif

// ---- error plumbing: every C-ABI entry point returns 0 or a cudaError_t value --------------
extern "C" void p2r_set_last_error(const char* where, int code);

#define P2R_RETURN_LAUNCH(where)                              \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) p2r_set_last_error(where, (int)_e);\
    return (int)_e;                                           \
  } while (0)

#define P2R_CHECK_ARG(cond, where)                            \
  do {                                                        \
    if (!(cond)) { p2r_set_last_error(where ": bad argument: " #cond, -1); return -1; } \
  } while (0)

static inline int p2r_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Kernel launch and dynamic shared memory, spelled so that a kernel that needs nothing beyond threads, shared memory,
// __syncthreads, shuffles and global atomics can ALSO be compiled for the host emulator of the CPU tests (launcher
// included: its grid / workspace arithmetic is then exercised too).  `kernel` must be a plain identifier (take a
// function pointer to a template instance first).
#ifdef P2R_HOST_EMULATION
#define P2R_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu_launch(dim3(grid), dim3(block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })
#define P2R_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu_dyn_smem)
#define P2R_DYN_SMEM_ALIGNED(type, name, align) type* name = reinterpret_cast<type*>(emu_dyn_smem)
#define P2R_NAMED_BARRIER_SYNC_1_256() emu_named_barrier_256()
#else
#define P2R_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define P2R_DYN_SMEM(type, name)                                \
  extern __shared__ __align__(16) unsigned char name##_raw[];   \
  type* name = reinterpret_cast<type*>(name##_raw)
#define P2R_DYN_SMEM_ALIGNED(type, name, align)                    \
  extern __shared__ __align__(align) unsigned char name##_raw[];   \
  type* name = reinterpret_cast<type*>(name##_raw)
// named barrier 1 over the 256 consumer threads of the streaming kernels (the producer warp does not take part)
#define P2R_NAMED_BARRIER_SYNC_1_256() asm volatile("bar.sync 1, 256;" ::: "memory")
#endif

// ---- exact-order fp32 arithmetic -----------------------------------------------------------
// The reference's kernels are compiled with nvcc's default -fmad=true; its SASS for sm_100a
// evaluates dx*dx + dy*dy + dz*dz as  t = dy*dy (FMUL); t = fma(dx,dx,t); t = fma(dz,dz,t)
// (see oracle/pointnet2_ref.c header).  The intrinsics pin that order regardless of how this
// translation unit is optimised, so distances are bit-identical with the reference kernels.
__device__ __forceinline__ float p2r_sqnorm3(float x, float y, float z) {
  float t = __fmul_rn(y, y);
  t = __fmaf_rn(x, x, t);
  t = __fmaf_rn(z, z, t);
  return t;
}
__device__ __forceinline__ float p2r_sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  return p2r_sqnorm3(__fsub_rn(ax, bx), __fsub_rn(ay, by), __fsub_rn(az, bz));
}
// torch's CPU vector norm over 3 components: sqrt(fma(z,z, fma(y,y, x*x))) (measured)
__device__ __forceinline__ float p2r_sqnorm3_xyz(float x, float y, float z) {
  float t = __fmul_rn(x, x);
  t = __fmaf_rn(y, y, t);
  t = __fmaf_rn(z, z, t);
  return t;
}

#ifndef P2R_HOST_EMULATION   // everything below is device-only (PTX)

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) ---------------------------------
__device__ __forceinline__ uint32_t p2r_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void p2r_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p2r_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void p2r_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void p2r_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void p2r_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p2r_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void p2r_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p2r_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void p2r_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(p2r_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void p2r_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   p2r_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(p2r_smem_u32(bar))
               : "memory");
}

// Stage `nfloats` contiguous floats from global into shared memory.  Uses one bulk-TMA copy when
// the source is 16-byte aligned (the tail that is not a multiple of 4 floats is copied by
// threads), otherwise plain coalesced loads.  All threads of the CTA must call it; it ends with
// the data visible to every thread.  `bar` must have been initialised with count 1; `parity` is
// the phase the caller tracks (flip after each use).
__device__ __forceinline__ void p2r_stage_floats(float* dst, const float* __restrict__ src, int nfloats,
                                                 uint64_t* bar, uint32_t parity) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int bulk = aligned ? (nfloats & ~3) : 0;
  if (threadIdx.x == 0) {
    if (bulk > 0) {
      p2r_mbar_expect_tx(bar, (uint32_t)bulk * 4u);
      p2r_bulk_g2s(dst, src, (uint32_t)bulk * 4u, bar);
    } else {
      p2r_mbar_arrive(bar);  // keep the phase bookkeeping uniform for the caller
    }
  }
  for (int i = bulk + threadIdx.x; i < nfloats; i += blockDim.x) dst[i] = __ldg(src + i);
  p2r_mbar_wait(bar, parity);
  __syncthreads();
}

#else   // host emulation: mbarrier / bulk-copy twins come from tests/csrc/cuda_emu.h; the staging helper is the same code
__device__ __forceinline__ void p2r_stage_floats(float* dst, const float* __restrict__ src, int nfloats,
                                                 uint64_t* bar, uint32_t parity) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  const int bulk = aligned ? (nfloats & ~3) : 0;
  if (threadIdx.x == 0) {
    if (bulk > 0) {
      p2r_mbar_expect_tx(bar, (uint32_t)bulk * 4u);
      p2r_bulk_g2s(dst, src, (uint32_t)bulk * 4u, bar);
    } else {
      p2r_mbar_arrive(bar);
    }
  }
  for (int i = bulk + threadIdx.x; i < nfloats; i += blockDim.x) dst[i] = __ldg(src + i);
  p2r_mbar_wait(bar, parity);
  __syncthreads();
}
#endif  // !P2R_HOST_EMULATION
