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
#include <string.h>

#include <barrier>
#include <memory>
#include <thread>
#include <vector>

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
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
  std::barrier<> bar{32};
};

static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;
static thread_local emu_warp* emu_my_warp = nullptr;
static thread_local int emu_lane = 0;
static std::barrier<>* emu_cta_barrier = nullptr;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
// dynamic shared memory (common.cuh's SCAE_DYNAMIC_SMEM): one host buffer, CTAs run one after another
alignas(16) static float emu_dynamic_smem[64 * 1024];
#define SCAE_DYNAMIC_SMEM(name) float* name = emu_dynamic_smem
template <class T>
inline T min(T a, T b) {
  return b < a ? b : a;
}

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
#define __align__(n) alignas(n)
inline float __frcp_rn(float x) { return 1.0f / x; }
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

// called by every emulated thread when its kernel body returns (ptx_emu.h flushes pending bulk stores there)
static void (*emu_thread_exit_hook)() = nullptr;

// kernel<<<grid, block>>>(args...) -> emu_launch(grid, block, [&] { kernel(args...); })
template <class F>
void emu_launch(int grid, int block, F body) {
  gridDim.x = (unsigned)grid;
  blockDim.x = (unsigned)block;
  for (int cta = 0; cta < grid; ++cta) {
    std::barrier<> cta_bar(block);
    emu_cta_barrier = &cta_bar;
    std::vector<std::unique_ptr<emu_warp>> warps;
    for (int w = 0; w < (block + 31) / 32; ++w) warps.emplace_back(new emu_warp);
    std::vector<std::thread> threads;
    for (int t = 0; t < block; ++t) {
      threads.emplace_back([&, t] {
        threadIdx.x = (unsigned)t;
        blockIdx.x = (unsigned)cta;
        emu_my_warp = warps[t / 32].get();
        emu_lane = t % 32;
        body();
        if (emu_thread_exit_hook) emu_thread_exit_hook();
      });
    }
    for (auto& th : threads) th.join();
  }
}
