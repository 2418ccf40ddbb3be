// Host stand-ins for the inline-PTX primitives of torch_scae_b200/csrc/ptx_sm100.cuh (TEST INFRASTRUCTURE): same names,
// same arguments, so that the capsule kernels' device code runs unmodified on the CPU under tests/emu/simt.h.
//
// The asynchronous copies are emulated as DEFERRED work, which is what makes ordering bugs visible:
//   * a bulk global->shared copy is only queued when issued; the bytes land when some thread WAITS on the mbarrier
//     (and the phase flips once the expected byte count has landed and the arrival count is met) -- code that reads the
//     destination before waiting sees stale shared memory;
//   * a bulk shared->global store is only queued when issued; its source is read when the issuing thread executes
//     cp.async.bulk.wait_group.read (or exits) -- code that overwrites the source before that stores the wrong data.
#pragma once
#define SCAE_PTX_SM100_CUH_   // keeps the real header out

#include <barrier>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "simt.h"

namespace scae {

inline float ex2_approx(float x) { return exp2f(x); }
inline float lg2_approx(float x) { return log2f(x); }
inline float rcp_approx(float x) { return 1.0f / x; }
inline float sin_approx(float x) { return sinf(x); }
inline float cos_approx(float x) { return cosf(x); }

// REDUX over the lanes of `mask`: like the shuffles, every live lane of the warp publishes between two warp barriers
// (the kernels call these from converged code only)
inline unsigned emu_redux(unsigned mask, unsigned v, bool want_max) {
  emu_my_warp->slot[emu_lane] = v;
  emu_my_warp->bar.arrive_and_wait();
  unsigned r = want_max ? 0u : 0xffffffffu;
  for (int l = 0; l < 32; ++l)
    if (mask >> l & 1u) r = want_max ? std::max(r, emu_my_warp->slot[l]) : std::min(r, emu_my_warp->slot[l]);
  emu_my_warp->bar.arrive_and_wait();
  return r;
}
inline unsigned redux_max_u32(unsigned mask, unsigned v) { return emu_redux(mask, v, true); }
inline unsigned redux_min_u32(unsigned mask, unsigned v) { return emu_redux(mask, v, false); }

// bar.sync id, n: one reusable std::barrier per id, created by the first thread that arrives; reset for every CTA
inline std::mutex emu_named_mutex;
inline std::unique_ptr<std::barrier<>> emu_named_barriers[16];
inline void named_bar_sync(unsigned id, unsigned n_threads) {
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> lock(emu_named_mutex);
    if (!emu_named_barriers[id]) emu_named_barriers[id].reset(new std::barrier<>((std::ptrdiff_t)n_threads));
    b = emu_named_barriers[id].get();
  }
  b->arrive_and_wait();
}

struct emu_copy {
  void* dst;
  const void* src;
  unsigned bytes;
};
struct emu_mbarrier {
  int count = 0, pending = 0, phase = 0;
  long tx = 0;
  std::vector<emu_copy> queued;
};
inline std::mutex emu_async_mutex;
inline std::map<unsigned, emu_mbarrier> emu_mbarriers;            // keyed by shared-memory byte offset
inline thread_local std::vector<emu_copy> emu_pending_stores;     // bulk groups are per thread

// a phase completes as soon as every expected arrival has happened and every expected byte has landed
inline void emu_try_complete(emu_mbarrier& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1;
    b.pending = b.count;
  }
}
inline unsigned smem_u32(const void* p) {
  return (unsigned)((const char*)p - (const char*)emu_dynamic_smem);
}
// shared memory by 32-bit address = byte offset into the emulated dynamic shared memory
inline float lds_f32(unsigned addr) { return *reinterpret_cast<const float*>((const char*)emu_dynamic_smem + addr); }
inline unsigned lds_u32(unsigned addr) { return *reinterpret_cast<const unsigned*>((const char*)emu_dynamic_smem + addr); }
inline float4 lds_f32x4(unsigned addr) { return *reinterpret_cast<const float4*>((const char*)emu_dynamic_smem + addr); }
inline void sts_f32(unsigned addr, float v) { *reinterpret_cast<float*>((char*)emu_dynamic_smem + addr) = v; }
inline void sts_u32(unsigned addr, unsigned v) { *reinterpret_cast<unsigned*>((char*)emu_dynamic_smem + addr) = v; }
inline float2 lds_f32x2(unsigned addr) { return *reinterpret_cast<const float2*>((const char*)emu_dynamic_smem + addr); }
inline void lds_pred_f32(unsigned addr, float& a, bool pred) {
  if (pred) a = lds_f32(addr);
}
inline void lds_pred_f32x2(unsigned addr, float& a, float& b, bool pred) {
  const float* p = reinterpret_cast<const float*>((const char*)emu_dynamic_smem + addr);
  if (pred) a = p[0], b = p[1];
}
inline void lds_pred_f32x4(unsigned addr, float& a, float& b, float& c, float& d, bool pred) {
  const float* p = reinterpret_cast<const float*>((const char*)emu_dynamic_smem + addr);
  if (pred) a = p[0], b = p[1], c = p[2], d = p[3];
}
inline void sts_pred_u32(unsigned addr, unsigned v, bool pred) {
  if (pred) sts_u32(addr, v);
}
inline void sts_pred_f32x4(unsigned addr, float a, float b, float c, float d, bool pred) {
  float* p = reinterpret_cast<float*>((char*)emu_dynamic_smem + addr);
  if (pred) p[0] = a, p[1] = b, p[2] = c, p[3] = d;
}
inline void smem_add_pred_f32(unsigned addr, float a, bool pred) {
  if (pred) *reinterpret_cast<float*>((char*)emu_dynamic_smem + addr) += a;
}
inline void smem_add_pred_f32x2(unsigned addr, float a, float b, bool pred) {
  float* p = reinterpret_cast<float*>((char*)emu_dynamic_smem + addr);
  if (pred) p[0] += a, p[1] += b;
}
inline void smem_add_pred_f32x4(unsigned addr, float a, float b, float c, float d, bool pred) {
  float* p = reinterpret_cast<float*>((char*)emu_dynamic_smem + addr);
  if (pred) p[0] += a, p[1] += b, p[2] += c, p[3] += d;
}
inline void mbar_init(unsigned bar, unsigned count) {
  std::lock_guard<std::mutex> lock(emu_async_mutex);
  emu_mbarrier& b = emu_mbarriers[bar];
  b = emu_mbarrier();
  b.count = b.pending = (int)count;
}
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx(unsigned bar, unsigned bytes) {      // arrive + expect_tx
  std::lock_guard<std::mutex> lock(emu_async_mutex);
  emu_mbarrier& b = emu_mbarriers.at(bar);
  b.tx += bytes;
  b.pending -= 1;
  emu_try_complete(b);   // (expect_tx(0): nothing to wait for)
}
inline void mbar_arrive(unsigned bar) {
  std::lock_guard<std::mutex> lock(emu_async_mutex);
  emu_mbarrier& b = emu_mbarriers.at(bar);
  b.pending -= 1;
  emu_try_complete(b);
}
inline void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned bar) {
  if ((reinterpret_cast<uintptr_t>(smem_dst) | reinterpret_cast<uintptr_t>(gsrc) | bytes) & 15u) {
    fprintf(stderr, "emu: bulk_g2s operands must be 16-byte aligned / sized\n");
    abort();
  }
  std::lock_guard<std::mutex> lock(emu_async_mutex);
  emu_mbarriers.at(bar).queued.push_back({smem_dst, gsrc, bytes});
}
inline void mbar_wait(unsigned bar, unsigned parity) {
  for (;;) {
    {
      std::lock_guard<std::mutex> lock(emu_async_mutex);
      emu_mbarrier& b = emu_mbarriers.at(bar);
      if ((unsigned)b.phase != parity) return;                  // the phase with this parity has completed
      if (!b.queued.empty()) {
        for (const emu_copy& c : b.queued) {                    // the queued copies land now
          memcpy(c.dst, c.src, c.bytes);
          b.tx -= c.bytes;
        }
        b.queued.clear();
        emu_try_complete(b);
        if ((unsigned)b.phase != parity) return;
      }
    }
    std::this_thread::yield();
  }
}
inline bool mbar_test(unsigned bar, unsigned parity) {
  std::lock_guard<std::mutex> lock(emu_async_mutex);
  emu_mbarrier& b = emu_mbarriers.at(bar);
  if ((unsigned)b.phase != parity) return true;
  if (!b.queued.empty()) {                                      // like a wait, a test lets the queued copies land
    for (const emu_copy& c : b.queued) {
      memcpy(c.dst, c.src, c.bytes);
      b.tx -= c.bytes;
    }
    b.queued.clear();
    emu_try_complete(b);
  }
  return (unsigned)b.phase != parity;
}
inline void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
  if ((reinterpret_cast<uintptr_t>(gdst) | reinterpret_cast<uintptr_t>(smem_src) | bytes) & 15u) {
    fprintf(stderr, "emu: bulk_s2g operands must be 16-byte aligned / sized\n");
    abort();
  }
  emu_pending_stores.push_back({gdst, smem_src, bytes});
}
inline void bulk_commit() {}
inline void bulk_wait_read_all() {
  for (const emu_copy& c : emu_pending_stores) memcpy(c.dst, c.src, c.bytes);
  emu_pending_stores.clear();
}
inline void bulk_wait_all() { bulk_wait_read_all(); }
inline void emu_flush_bulk_stores_at_exit() { bulk_wait_read_all(); }

// cp.async: 4-byte copies queued per thread in groups; a group lands when cp_async_wait<N> leaves at most N younger
// groups pending (or when the thread exits)
inline thread_local std::vector<std::vector<emu_copy>> emu_cp_groups;
inline thread_local std::vector<emu_copy> emu_cp_open;
inline void cp_async4(float* smem_dst, const float* gsrc) { emu_cp_open.push_back({smem_dst, gsrc, 4u}); }
inline void cp_async_commit() {
  emu_cp_groups.push_back(emu_cp_open);
  emu_cp_open.clear();
}
inline void emu_cp_land(size_t keep_groups) {
  while (emu_cp_groups.size() > keep_groups) {
    for (const emu_copy& c : emu_cp_groups.front()) memcpy(c.dst, c.src, c.bytes);
    emu_cp_groups.erase(emu_cp_groups.begin());
  }
}
template <int N>
inline void cp_async_wait() {
  emu_cp_land((size_t)N);
}
inline void emu_async_thread_exit() {
  bulk_wait_read_all();
  cp_async_commit();
  emu_cp_land(0);
}

// every emulated thread completes its outstanding asynchronous copies when its kernel body returns
inline const bool emu_async_hook_installed = (emu_thread_exit_hook = emu_async_thread_exit, true);
// named barriers and mbarriers belong to one CTA
inline void emu_async_cta_start() {
  for (auto& b : emu_named_barriers) b.reset();
  emu_mbarriers.clear();
}
inline const bool emu_cta_hook_installed = (emu_cta_start_hook = emu_async_cta_start, true);

// optimisation barrier: nothing to hide from on the host
inline unsigned keep(unsigned v) { return v; }
inline float keep(float v) { return v; }
inline int keep(int v) { return v; }
inline size_t keep(size_t v) { return v; }

}  // namespace scae
