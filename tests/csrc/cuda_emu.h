// cuda_emu.h -- TEST HARNESS (CPU): a minimal CUDA execution-model emulator, enough to run the SIMT kernels of
// pose2room_b200/csrc that use nothing beyond threads / blocks, static + dynamic shared memory, __syncthreads, warp
// shuffles, global atomics and __threadfence (loss_ops.cu, gmm_ops.cu) on the host, so that their PLUMBING -- block
// reductions, the last-block-finalises pattern, arg-min merges, indexing, the launchers' grid / workspace arithmetic --
// is exercised without a GPU (tests/test_kernels_emulated.py).  Not a performance model and not part of the product:
// compiled only by the test (g++ -DP2R_HOST_EMULATION), never shipped, never loaded by pose2room_b200/.
//
// Execution: blocks run one after the other in a SHUFFLED order (the last block to finish is not always the last index);
// the threads of a block are real pthreads, __syncthreads is a pthread barrier over the block, a warp shuffle is an
// exchange through a per-warp slot array between two per-warp barriers.  `__shared__` becomes `static` (blocks are
// sequential, so one copy is right); dynamic shared memory is one buffer per launch.
#pragma once
#include <pthread.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <random>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
static thread_local uint3_emu threadIdx, blockIdx;
static dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }

// ---- block / warp state of the running launch -------------------------------------------------------------------
struct EmuWarp {
  pthread_barrier_t bar;
  unsigned char slot[32][16];
};
static pthread_barrier_t emu_block_bar;
static std::vector<EmuWarp>* emu_warps = nullptr;
static unsigned char* emu_dyn_smem = nullptr;
static unsigned long long emu_launches = 0;

static inline void __syncthreads() { pthread_barrier_wait(&emu_block_bar); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int off) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  EmuWarp& w = (*emu_warps)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  memcpy(w.slot[lane], &v, sizeof(T));
  pthread_barrier_wait(&w.bar);
  T r = v;
  if (lane + off < 32) memcpy(&r, w.slot[lane + off], sizeof(T));
  pthread_barrier_wait(&w.bar);
  return r;
}

template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  EmuWarp& w = (*emu_warps)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  memcpy(w.slot[lane], &v, sizeof(T));
  pthread_barrier_wait(&w.bar);
  T r;
  memcpy(&r, w.slot[src & 31], sizeof(T));
  pthread_barrier_wait(&w.bar);
  return r;
}

static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

// bf16 storage type: enough for the template specialisations to compile and round-trip
struct __nv_bfloat16 { unsigned short x; };
static inline float __bfloat162float(__nv_bfloat16 h) { unsigned u = (unsigned)h.x << 16; float f; memcpy(&f, &u, 4); return f; }
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  unsigned u; memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);              // round to nearest even (finite inputs)
  __nv_bfloat16 h; h.x = (unsigned short)(u >> 16); return h;
}

// ---- launch ---------------------------------------------------------------------------------------------------------
struct EmuThreadArg { const std::function<void()>* fn; unsigned tid, bid; };
static void* emu_thread_main(void* p) {
  EmuThreadArg* a = (EmuThreadArg*)p;
  threadIdx.x = a->tid; threadIdx.y = threadIdx.z = 0;
  blockIdx.x = a->bid; blockIdx.y = blockIdx.z = 0;
  (*a->fn)();
  return nullptr;
}

// 1-D grids and blocks (all the emulated kernels use); block size a multiple of 32
static inline void emu_launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& fn) {
  if (block % 32 != 0 || block == 0 || grid == 0) abort();
  gridDim = dim3(grid); blockDim = dim3(block);
  std::vector<unsigned char> dyn(smem + 16);
  emu_dyn_smem = dyn.data() + ((16 - ((uintptr_t)dyn.data() & 15)) & 15);
  std::vector<unsigned> order(grid);
  for (unsigned i = 0; i < grid; ++i) order[i] = i;
  std::mt19937 rng(12345u + (unsigned)(emu_launches++));
  std::shuffle(order.begin(), order.end(), rng);
  std::vector<EmuWarp> warps(block / 32);
  emu_warps = &warps;
  std::vector<pthread_t> th(block);
  std::vector<EmuThreadArg> args(block);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 256 * 1024);
  for (unsigned b : order) {
    pthread_barrier_init(&emu_block_bar, nullptr, block);
    for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    for (unsigned t = 0; t < block; ++t) {
      args[t] = EmuThreadArg{&fn, t, b};
      if (pthread_create(&th[t], &attr, emu_thread_main, &args[t]) != 0) abort();
    }
    for (unsigned t = 0; t < block; ++t) pthread_join(th[t], nullptr);
    pthread_barrier_destroy(&emu_block_bar);
    for (auto& w : warps) pthread_barrier_destroy(&w.bar);
  }
  pthread_attr_destroy(&attr);
  emu_warps = nullptr;
  emu_dyn_smem = nullptr;
}
