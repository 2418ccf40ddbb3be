// Host-side SIMT emulation for CPU tests of simple CUDA kernels (TEST INFRASTRUCTURE, never shipped).
//
// One std::thread per CUDA thread, one CTA at a time; __syncthreads() is a CTA-wide barrier and the *_sync warp
// shuffles exchange values through a per-warp buffer between two warp-wide barriers, so divergence bugs (a shuffle or
// barrier not reached by every thread) dead-lock here just as they would hang the GPU.  __shared__ variables become
// function-local statics (shared by all threads; CTAs run one after another).  Only what the tested kernels use is
// provided (csrc/loss_head.cu, csrc/attnpool_cl.cu, csrc/conv_cols.cu): no textures, atomics, TMA or asynchronous copies.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdint.h>

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

struct emu_warp {
  float slot[32];
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
static float emu_dynamic_smem[64 * 1024];
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
inline float __shfl_xor_sync(unsigned, float v, int d) {
  emu_my_warp->slot[emu_lane] = v;
  emu_my_warp->bar.arrive_and_wait();
  const float r = emu_my_warp->slot[emu_lane ^ d];
  emu_my_warp->bar.arrive_and_wait();
  return r;
}

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
      });
    }
    for (auto& th : threads) th.join();
  }
}
