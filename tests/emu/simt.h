// Host-side SIMT emulation for CPU tests of simple CUDA kernels (TEST INFRASTRUCTURE, never shipped).
//
// One std::thread per CUDA thread, one CTA at a time; __syncthreads() is a CTA-wide barrier and the *_sync warp
// shuffles exchange values through a per-warp buffer between two warp-wide barriers, so divergence bugs (a shuffle or
// barrier not reached by every thread) dead-lock here just as they would hang the GPU.  __shared__ variables become
// function-local statics (shared by all threads; CTAs run one after another).  Only what the tested kernels use is
// provided (csrc/loss_head.cu, csrc/attnpool_cl.cu, csrc/conv_cols.cu, csrc/caps_ll2.cu with tests/emu/ptx_emu.h): no textures, atomics, TMA or asynchronous copies.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct float2 {
  float x, y;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }

struct emu_warp {
  uint32_t slot[32];
  std::barrier<> bar;
  explicit emu_warp(int lanes) : bar(lanes) {}
};

// (inline variables: ONE instance per program / shared library -- the csrc headers define inline device functions such as
// warp_sum that several translation units share, so per-file statics would split the state)
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
inline thread_local emu_warp* emu_my_warp = nullptr;
inline thread_local int emu_lane = 0;
inline std::barrier<>* emu_cta_barrier = nullptr;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
// dynamic shared memory (common.cuh's SCAE_DYNAMIC_SMEM): one host buffer, CTAs run one after another
constexpr size_t kEmuSmemFloats = 64 * 1024;
alignas(16) inline float emu_smem_storage[kEmuSmemFloats];
inline float* emu_dynamic_smem = emu_smem_storage;   // (with -DEMU_EXACT_SMEM: a heap block of exactly the launch's size)
#define SCAE_DYNAMIC_SMEM(name) float* name = emu_dynamic_smem
template <class T>
inline T min(T a, T b) {
  return b < a ? b : a;
}
template <class T>
inline T max(T a, T b) {
  return a < b ? b : a;
}
inline long min(long a, int b) { return b < a ? (long)b : a; }
inline long min(int a, long b) { return b < a ? b : (long)a; }
inline long max(long a, int b) { return a < b ? (long)b : a; }
inline long max(int a, long b) { return a < b ? b : (long)a; }

template <class T>
inline T __ldg(const T* p) {
  return *p;
}
inline void __syncthreads() { emu_cta_barrier->arrive_and_wait(); }
inline void __syncwarp() { emu_my_warp->bar.arrive_and_wait(); }
// warp shuffles for 4-byte types: every lane publishes its value between two warp-wide barriers
template <class T>
inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  uint32_t bits;
  memcpy(&bits, &v, 4);
  emu_my_warp->slot[emu_lane] = bits;
  emu_my_warp->bar.arrive_and_wait();
  bits = emu_my_warp->slot[src_lane & 31];
  emu_my_warp->bar.arrive_and_wait();
  T r;
  memcpy(&r, &bits, 4);
  return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int d) {
  return emu_shfl(v, emu_lane ^ d);
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
  return emu_shfl(v, src);
}

// device math intrinsics used by the csrc headers
#define __align__(n) __attribute__((aligned(n)))
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float x) { return sqrtf(x); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline void sincospif(float x, float* s, float* c) {
  *s = (float)sin(M_PI * (double)x);
  *c = (float)cos(M_PI * (double)x);
}
inline int __float_as_int(float f) {
  int i;
  memcpy(&i, &f, 4);
  return i;
}
inline float __int_as_float(int i) {
  float f;
  memcpy(&f, &i, 4);
  return f;
}

inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline float __fadd_rn(float a, float b) { return a + b; }      // (compiled with -ffp-contract=off: never fused)
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fadd_rd(float a, float b) {                       // round toward -infinity
  const double exact = (double)a + (double)b;                    // exact: both addends are floats
  float r = (float)exact;
  if ((double)r > exact) r = nextafterf(r, -INFINITY);
  return r;
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  return emu_shfl(v, emu_lane >= (int)delta ? emu_lane - (int)delta : emu_lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  emu_my_warp->slot[emu_lane] = pred ? 1u : 0u;
  emu_my_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (emu_my_warp->slot[l] & 1u) << l;
  emu_my_warp->bar.arrive_and_wait();
  return m;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned __match_any_sync(unsigned, unsigned value) {
  emu_my_warp->slot[emu_lane] = value;
  emu_my_warp->bar.arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (emu_my_warp->slot[l] == value ? 1u : 0u) << l;
  emu_my_warp->bar.arrive_and_wait();
  return m;
}

// called by every emulated thread when its kernel body returns (ptx_emu.h flushes pending bulk stores there)
inline void (*emu_thread_exit_hook)() = nullptr;
// called once before the threads of every emulated CTA start (ptx_emu.h resets named barriers / mbarriers there)
inline void (*emu_cta_start_hook)() = nullptr;

// kernel<<<grid, block, smem, stream>>>(args...) -> emu_launch(grid, block, [&] { kernel(args...); })
// (tests/emu/build_lib.py rewrites the launches of whole .cu files this way).  A thread whose body returns leaves the
// CTA and warp barriers, so kernels with early exits do not dead-lock the threads that go on.
template <class F>
void emu_launch(dim3 grid, dim3 block, F body, size_t smem_bytes = (size_t)-1) {
  gridDim = grid;
  blockDim = block;
  const int n_threads = (int)block.x;
#ifdef EMU_EXACT_SMEM
  // memcheck mode (build with AddressSanitizer): dynamic shared memory is a heap block of exactly the size the launch
  // asked for, so an overrun of the kernel's shared-memory layout lands in a redzone
  struct Exact {
    float* saved = emu_dynamic_smem;
    void* block = nullptr;
    explicit Exact(size_t bytes) {
      if (bytes != (size_t)-1) {
        block = aligned_alloc(16, (bytes + 15) / 16 * 16 + 16);
        emu_dynamic_smem = static_cast<float*>(block);
      }
    }
    ~Exact() {
      emu_dynamic_smem = saved;
      free(block);
    }
  } exact(smem_bytes);
#else
  if (smem_bytes != (size_t)-1 && smem_bytes > sizeof(emu_smem_storage)) {
    fprintf(stderr, "emu_launch: %zu bytes of dynamic shared memory requested, %zu available\n", smem_bytes,
            sizeof(emu_smem_storage));
    abort();
  }
#endif
  for (unsigned by = 0; by < grid.y; ++by) {
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      std::barrier<> cta_bar(n_threads);
      emu_cta_barrier = &cta_bar;
      if (emu_cta_start_hook) emu_cta_start_hook();
      std::vector<std::unique_ptr<emu_warp>> warps;
      for (int w = 0; w * 32 < n_threads; ++w) warps.emplace_back(new emu_warp(std::min(32, n_threads - w * 32)));
      for (auto& w : warps)
        for (int l = 0; l < 32; ++l) w->slot[l] = 0;
      std::vector<std::thread> threads;
      for (int t = 0; t < n_threads; ++t) {
        threads.emplace_back([&, t] {
          threadIdx = dim3((unsigned)t);
          blockIdx = dim3(bx, by);
          emu_my_warp = warps[t / 32].get();
          emu_lane = t % 32;
          body();
          if (emu_thread_exit_hook) emu_thread_exit_hook();
          emu_my_warp->bar.arrive_and_drop();
          cta_bar.arrive_and_drop();
        });
      }
      for (auto& th : threads) th.join();
    }
  }
}
template <class F>
void emu_launch(int grid, int block, F body) {
  emu_launch(dim3((unsigned)grid), dim3((unsigned)block), body);
}
