// cuda_emu.h -- TEST HARNESS (CPU): a minimal CUDA execution-model emulator, enough to run the SIMT kernels of
// pose2room_b200/csrc that use nothing beyond threads / blocks, static + dynamic shared memory, __syncthreads, warp
// shuffles / votes, global atomics and __threadfence (loss_ops.cu, gmm_ops.cu, vote_ops.cu, geometry_ops.cu) on the
// host, so that their PLUMBING -- block reductions, the last-block-finalises pattern, arg-min merges, indexing, tail
// blocks, the launchers' grid / workspace arithmetic -- is exercised without a GPU (tests/test_kernels_emulated.py,
// tests/test_eval_kernels_emulated.py).  Not a performance model and not part of the product: compiled only by the
// tests (g++ -DP2R_HOST_EMULATION), never shipped, never loaded by pose2room_b200/.
//
// Execution: one pool of blockDim.x pthreads per launch walks over the blocks of the grid one after the other, in a
// SHUFFLED order (the last block to finish is not always the last index).  __syncthreads is a barrier over the threads
// of the block that have not returned from the kernel yet (like the hardware's), a warp shuffle / vote is an exchange
// through a per-warp slot array between two per-warp barriers.  `__shared__` becomes `static` (blocks are sequential, so
// one copy is right); dynamic shared memory is one buffer per launch.  Floating-point intrinsics map to the plain
// operation: compile with -ffp-contract=off so that nothing is fused behind their back.
#pragma once
#include <pthread.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <map>
#include <stdio.h>
#include <time.h>
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
#define __align__(n) __attribute__((aligned(n)))

struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct __attribute__((aligned(8))) float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }

// ---- a barrier whose participants may leave (a thread that returns from the kernel stops counting) ----------------
struct EmuBarrier {
  pthread_mutex_t m;
  pthread_cond_t cv;
  int count, waiting;
  unsigned gen;
  EmuBarrier() : count(0), waiting(0), gen(0) { pthread_mutex_init(&m, nullptr); pthread_cond_init(&cv, nullptr); }
  ~EmuBarrier() { pthread_mutex_destroy(&m); pthread_cond_destroy(&cv); }
  void reset(int n) { count = n; waiting = 0; }
  void wait() {
    pthread_mutex_lock(&m);
    if (++waiting >= count) {
      waiting = 0; ++gen;
      pthread_cond_broadcast(&cv);
    } else {
      const unsigned g = gen;
      while (g == gen) pthread_cond_wait(&cv, &m);
    }
    pthread_mutex_unlock(&m);
  }
  void leave() {
    pthread_mutex_lock(&m);
    --count;
    if (count > 0 && waiting >= count) {
      waiting = 0; ++gen;
      pthread_cond_broadcast(&cv);
    }
    pthread_mutex_unlock(&m);
  }
};

struct EmuWarp {
  EmuBarrier bar;
  unsigned char slot[32][16];
};
static EmuBarrier* emu_block_bar = nullptr;
static std::vector<EmuWarp>* emu_warps = nullptr;
static unsigned char* emu_dyn_smem = nullptr;
static unsigned long long emu_launches = 0;

static inline void __syncthreads() { emu_block_bar->wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { (*emu_warps)[threadIdx.x >> 5].bar.wait(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// every lane publishes v, reads the slot of lane `src` (its own value when src is outside the warp)
template <typename T>
static inline T emu_exchange(T v, int src) {
  static_assert(sizeof(T) <= 16, "shuffle payload");
  EmuWarp& w = (*emu_warps)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  memcpy(w.slot[lane], &v, sizeof(T));
  w.bar.wait();
  T r = v;
  if (src >= 0 && src < 32) memcpy(&r, w.slot[src], sizeof(T));
  w.bar.wait();
  return r;
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int off) { return emu_exchange(v, (int)(threadIdx.x & 31) + off); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) { return emu_exchange(v, (int)(threadIdx.x & 31) ^ mask); }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
static inline int __any_sync(unsigned, int pred) {
  EmuWarp& w = (*emu_warps)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31, p = pred != 0;
  memcpy(w.slot[lane], &p, sizeof(int));
  w.bar.wait();
  int any = 0;
  for (int l = 0; l < 32; ++l) { int q; memcpy(&q, w.slot[l], sizeof(int)); any |= q; }
  w.bar.wait();
  return any;
}

static inline unsigned __reduce_max_sync(unsigned, unsigned v) {      // redux.sync.max.u32
  EmuWarp& w = (*emu_warps)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  memcpy(w.slot[lane], &v, sizeof(unsigned));
  w.bar.wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) { unsigned q; memcpy(&q, w.slot[l], sizeof(unsigned)); m = q > m ? q : m; }
  w.bar.wait();
  return m;
}

static inline unsigned __ballot_sync(unsigned, int pred) {
  EmuWarp& w = (*emu_warps)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31, p = pred != 0;
  memcpy(w.slot[lane], &p, sizeof(int));
  w.bar.wait();
  unsigned bal = 0;
  for (int l = 0; l < 32; ++l) { int q; memcpy(&q, w.slot[l], sizeof(int)); bal |= (unsigned)(q != 0) << l; }
  w.bar.wait();
  return bal;
}
// block-wide AND of a predicate: and -> barrier -> read -> barrier -> re-arm -> barrier
static int emu_and_acc = 1;
static inline int __syncthreads_and(int pred) {
  if (!pred) __atomic_store_n(&emu_and_acc, 0, __ATOMIC_SEQ_CST);
  emu_block_bar->wait();
  const int r = __atomic_load_n(&emu_and_acc, __ATOMIC_SEQ_CST);
  emu_block_bar->wait();
  __atomic_store_n(&emu_and_acc, 1, __ATOMIC_SEQ_CST);
  emu_block_bar->wait();
  return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
  unsigned old, neu;
  float f;
  do {
    old = __atomic_load_n((unsigned*)p, __ATOMIC_SEQ_CST);
    memcpy(&f, &old, 4); f += v; memcpy(&neu, &f, 4);
  } while (!__atomic_compare_exchange_n((unsigned*)p, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  memcpy(&f, &old, 4);
  return f;
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
template <typename T>
static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

// round-to-nearest intrinsics = the plain operation (the build uses -ffp-contract=off); fma intrinsics = libm fma
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
template <typename T>
static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T>
static inline T max(T a, T b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
static inline long long min(int a, long long b) { return a < b ? (long long)a : b; }

// bf16 storage type: enough for the template specialisations to compile and round-trip
struct __nv_bfloat16 { unsigned short x; };
static inline float __bfloat162float(__nv_bfloat16 h) { unsigned u = (unsigned)h.x << 16; float f; memcpy(&f, &u, 4); return f; }
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  unsigned u; memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);              // round to nearest even (finite inputs)
  __nv_bfloat16 h; h.x = (unsigned short)(u >> 16); return h;
}


// ---- mbarrier + 1-D bulk copy (cp.async.bulk) twins of the helpers in p2r_common.cuh, a few vector types -------------
// An mbarrier is a side-table entry keyed by the address of the kernel's uint64_t: pending arrivals, outstanding
// transaction bytes, phase parity.  A bulk copy happens at issue time and then completes its bytes on the barrier.  A
// wait that makes no progress for 10 s aborts with a diagnostic (a deadlock in a ring protocol would otherwise hang).
struct uint4 { unsigned x, y, z, w; } __attribute__((aligned(16)));
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { __nv_bfloat162 r; r.x = __float2bfloat16_rn(a); r.y = __float2bfloat16_rn(b); return r; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline double atomicAdd(double* p, double v) {
  unsigned long long old, neu; double f;
  do { old = __atomic_load_n((unsigned long long*)p, __ATOMIC_SEQ_CST); memcpy(&f, &old, 8); f += v; memcpy(&neu, &f, 8);
  } while (!__atomic_compare_exchange_n((unsigned long long*)p, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  memcpy(&f, &old, 8); return f;
}
struct EmuMbar { int init, pending; long long tx; unsigned phase; };
static pthread_mutex_t emu_mbar_m = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t emu_mbar_cv = PTHREAD_COND_INITIALIZER;
static std::map<const void*, EmuMbar> emu_mbars;
static inline void emu_mbar_check(EmuMbar& b) {
  if (b.pending <= 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.init; pthread_cond_broadcast(&emu_mbar_cv); }
}
static inline void p2r_mbar_init(uint64_t* bar, uint32_t count) {
  pthread_mutex_lock(&emu_mbar_m); emu_mbars[bar] = EmuMbar{(int)count, (int)count, 0, 0u}; pthread_mutex_unlock(&emu_mbar_m);
}
static inline void p2r_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  pthread_mutex_lock(&emu_mbar_m); EmuMbar& b = emu_mbars[bar]; b.tx += bytes; b.pending -= 1; emu_mbar_check(b); pthread_mutex_unlock(&emu_mbar_m);
}
static inline void p2r_mbar_arrive(uint64_t* bar) {
  pthread_mutex_lock(&emu_mbar_m); EmuMbar& b = emu_mbars[bar]; b.pending -= 1; emu_mbar_check(b); pthread_mutex_unlock(&emu_mbar_m);
}
static inline void p2r_mbar_wait(uint64_t* bar, uint32_t parity) {
  pthread_mutex_lock(&emu_mbar_m);
  EmuMbar& b = emu_mbars[bar];
  struct timespec ts; int spins = 0;
  while (b.phase == parity) {
    clock_gettime(CLOCK_REALTIME, &ts); ts.tv_sec += 5;
    if (pthread_cond_timedwait(&emu_mbar_cv, &emu_mbar_m, &ts) != 0 && ++spins >= 2) {
      fprintf(stderr, "EMU: mbarrier wait stuck: block %u thread %u bar %p parity %u (pending %d tx %lld phase %u)\n",
              blockIdx.x, threadIdx.x, (void*)bar, parity, b.pending, b.tx, b.phase);
      abort();
    }
  }
  pthread_mutex_unlock(&emu_mbar_m);
}
static inline void p2r_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  memcpy(dst, src, bytes);
  pthread_mutex_lock(&emu_mbar_m); EmuMbar& b = emu_mbars[bar]; b.tx -= bytes; emu_mbar_check(b); pthread_mutex_unlock(&emu_mbar_m);
}
static EmuBarrier emu_named_256;
static inline void emu_named_barrier_256() { emu_named_256.wait(); }
static inline void p2r_fence_mbar_init() {}
static inline void p2r_fence_proxy_async() {}

// ---- launch ---------------------------------------------------------------------------------------------------------
struct EmuLaunch {
  const std::function<void()>* fn;
  std::vector<unsigned> order;       // linear block ids, shuffled
  pthread_barrier_t gate;            // all pool threads, between blocks
  EmuBarrier block_bar;
  std::vector<EmuWarp> warps;
  unsigned nthreads;
};
struct EmuThreadArg { EmuLaunch* L; unsigned tid; };

static void* emu_thread_main(void* p) {
  EmuThreadArg* a = (EmuThreadArg*)p;
  EmuLaunch& L = *a->L;
  threadIdx.x = a->tid; threadIdx.y = threadIdx.z = 0;
  for (unsigned lin : L.order) {
    if (a->tid == 0) {               // (the gate below orders this reset after every thread has left the previous block)
      L.block_bar.reset((int)L.nthreads);
      for (auto& w : L.warps) w.bar.reset(32);
    }
    pthread_barrier_wait(&L.gate);
    blockIdx.x = lin % gridDim.x;
    blockIdx.y = (lin / gridDim.x) % gridDim.y;
    blockIdx.z = lin / (gridDim.x * gridDim.y);
    (*L.fn)();
    L.warps[a->tid >> 5].bar.leave();
    L.block_bar.leave();
    pthread_barrier_wait(&L.gate);
  }
  return nullptr;
}

// any grid; 1-D blocks whose size is a multiple of 32
static inline void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& fn) {
  if (block.y != 1 || block.z != 1 || block.x % 32 != 0 || block.x == 0 || grid.x * grid.y * grid.z == 0) abort();
  gridDim = grid; blockDim = block;
  std::vector<unsigned char> dyn(smem + 16);
  emu_dyn_smem = dyn.data() + ((16 - ((uintptr_t)dyn.data() & 15)) & 15);
  EmuLaunch L;
  L.fn = &fn;
  L.nthreads = block.x;
  L.order.resize((size_t)grid.x * grid.y * grid.z);
  for (size_t i = 0; i < L.order.size(); ++i) L.order[i] = (unsigned)i;
  std::mt19937 rng(12345u + (unsigned)(emu_launches++));
  std::shuffle(L.order.begin(), L.order.end(), rng);
  L.warps = std::vector<EmuWarp>(block.x / 32);
  pthread_barrier_init(&L.gate, nullptr, block.x);
  emu_block_bar = &L.block_bar;
  emu_named_256.reset(256);
  emu_warps = &L.warps;
  std::vector<pthread_t> th(block.x);
  std::vector<EmuThreadArg> args(block.x);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 512 * 1024);
  for (unsigned t = 0; t < block.x; ++t) {
    args[t] = EmuThreadArg{&L, t};
    if (pthread_create(&th[t], &attr, emu_thread_main, &args[t]) != 0) abort();
  }
  for (unsigned t = 0; t < block.x; ++t) pthread_join(th[t], nullptr);
  pthread_attr_destroy(&attr);
  pthread_barrier_destroy(&L.gate);
  emu_block_bar = nullptr;
  emu_warps = nullptr;
  emu_dyn_smem = nullptr;
}
