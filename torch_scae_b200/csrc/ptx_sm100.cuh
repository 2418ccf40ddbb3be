// Every inline-PTX primitive of the likelihood kernels in one place (sm_100a): MUFU approximations, mbarrier and 1-D
// bulk (TMA) copies, cp.async, the optimisation barrier.  Kept apart from the arithmetic that uses them so that tests/emu can execute the kernels'
// device code on the CPU with host stand-ins for exactly these functions (tests/emu/ptx_emu.h) and nothing else.
#ifndef SCAE_PTX_SM100_CUH_   // (a classic guard: the emulation pre-defines it to substitute its stand-ins)
#define SCAE_PTX_SM100_CUH_

#include <stdint.h>

namespace scae {

// ---- MUFU approximations (2^x, log2 x, 1/x; flush-to-zero) ---------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// sin / cos of an angle in radians, |x| <= pi for full accuracy (MUFU.SIN / MUFU.COS after the hardware's 1/2pi scaling)
__device__ __forceinline__ float sin_approx(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float cos_approx(float x) {
  float y;
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- warp-level integer reductions over the lanes of `mask` (REDUX); every lane of `mask` must call with that mask ----
__device__ __forceinline__ unsigned redux_max_u32(unsigned mask, unsigned v) {
  unsigned r;
  asm volatile("redux.sync.max.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(mask));
  return r;
}
__device__ __forceinline__ unsigned redux_min_u32(unsigned mask, unsigned v) {
  unsigned r;
  asm volatile("redux.sync.min.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(mask));
  return r;
}

// ---- named barrier among `n_threads` (a multiple of 32) threads of the CTA; id 0 is __syncthreads' ---------------------
__device__ __forceinline__ void named_bar_sync(unsigned id, unsigned n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// ---- shared memory by 32-bit address (what smem_u32 returns): the persistent capsule kernels do their own address
// arithmetic so that a pair's accesses are one base register plus immediates -------------------------------------------
__device__ __forceinline__ float lds_f32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(unsigned addr) {   // 16-byte aligned
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(unsigned addr, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float2 lds_f32x2(unsigned addr) {   // 8-byte aligned
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
// Predicated loads: the destination keeps its value where `pred` does not hold.
__device__ __forceinline__ void lds_pred_f32(unsigned addr, float& a, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.f32 %0, [%1];\n\t}"
               : "+f"(a)
               : "r"(addr), "r"((unsigned)pred)
               : "memory");
}
__device__ __forceinline__ void lds_pred_f32x2(unsigned addr, float& a, float& b, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.v2.f32 {%0, %1}, [%2];\n\t}"
               : "+f"(a), "+f"(b)
               : "r"(addr), "r"((unsigned)pred)
               : "memory");
}
__device__ __forceinline__ void lds_pred_f32x4(unsigned addr, float& a, float& b, float& c, float& d, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+f"(a), "+f"(b), "+f"(c), "+f"(d)
               : "r"(addr), "r"((unsigned)pred)
               : "memory");
}
// Predicated stores.
__device__ __forceinline__ void sts_pred_u32(unsigned addr, unsigned v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v),
               "r"((unsigned)pred)
               : "memory");
}
__device__ __forceinline__ void sts_pred_f32x4(unsigned addr, float a, float b, float c, float d, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"r"(addr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"((unsigned)pred)
               : "memory");
}
// Predicated read-modify-write of 1 / 2 / 4 consecutive floats: [addr] += v where `pred` holds, as predicated
// instructions (no branch, no reconvergence bookkeeping).  Callers guarantee that the lanes with `pred` set use distinct
// addresses.
__device__ __forceinline__ void smem_add_pred_f32(unsigned addr, float a, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 t;\n\tsetp.ne.u32 p, %2, 0;\n\t"
      "@p ld.shared.f32 t, [%0];\n\t@p add.f32 t, t, %1;\n\t@p st.shared.f32 [%0], t;\n\t}" ::"r"(addr),
      "f"(a), "r"((unsigned)pred)
      : "memory");
}
__device__ __forceinline__ void smem_add_pred_f32x2(unsigned addr, float a, float b, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 t, u;\n\tsetp.ne.u32 p, %3, 0;\n\t"
      "@p ld.shared.v2.f32 {t, u}, [%0];\n\t@p add.f32 t, t, %1;\n\t@p add.f32 u, u, %2;\n\t"
      "@p st.shared.v2.f32 [%0], {t, u};\n\t}" ::"r"(addr),
      "f"(a), "f"(b), "r"((unsigned)pred)
      : "memory");
}
__device__ __forceinline__ void smem_add_pred_f32x4(unsigned addr, float a, float b, float c, float d, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .f32 t, u, v, w;\n\tsetp.ne.u32 p, %5, 0;\n\t"
      "@p ld.shared.v4.f32 {t, u, v, w}, [%0];\n\t@p add.f32 t, t, %1;\n\t@p add.f32 u, u, %2;\n\t"
      "@p add.f32 v, v, %3;\n\t@p add.f32 w, w, %4;\n\t@p st.shared.v4.f32 [%0], {t, u, v, w};\n\t}" ::"r"(addr),
      "f"(a), "f"(b), "f"(c), "f"(d), "r"((unsigned)pred)
      : "memory");
}

// ---- mbarrier + bulk (TMA) copies, 1-D --------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (bulk copies reading or overwriting them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// plain arrival (release at CTA scope): counts one of the barrier's expected arrivals
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// global -> shared, completion counted in bytes on `bar`; dst, src and bytes multiples of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global; src, dst and bytes multiples of 16
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk stores of this thread have COMPLETED (their global writes are performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- cp.async (Ampere-style 4-byte asynchronous copies; the general capsule path stages ragged rows with them) -------
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- optimisation barrier --------------------------------------------------------------------------------------------
// Hides a loop-invariant value from the optimiser: ptxas otherwise re-derives it from the kernel parameters in every
// pass of the hot loop (rematerialisation) instead of keeping it in a register.
__device__ __forceinline__ unsigned keep(unsigned v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ float keep(float v) {
  asm volatile("" : "+f"(v));
  return v;
}
__device__ __forceinline__ int keep(int v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ size_t keep(size_t v) {
  asm volatile("" : "+l"(v));
  return v;
}

}  // namespace scae
#endif  // SCAE_PTX_SM100_CUH_
