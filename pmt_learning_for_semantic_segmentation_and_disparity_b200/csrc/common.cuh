// common.cuh -- shared helpers for libpmt_ops (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pmt_ops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpmt_ops is written for sm_100a (B200) only"
#endif

namespace pmt {

// ---- error plumbing (thread-local message, integer codes across the C ABI) -------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define PMT_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::pmt::set_error(__VA_ARGS__);    \
      return PMT_ERR_INVALID;           \
    }                                   \
  } while (0)

#define PMT_CUDA_OK(expr)                                                                  \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::pmt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PMT_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define PMT_LAUNCH_OK(name)                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::pmt::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));           \
      return PMT_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

int sm_count();  // cached number of SMs of the current device

// Tuning / ablation knobs.  They exist only in development builds (-DPMT_DEV_KNOBS, `PMT_DEV_KNOBS=1` in the
// environment of _build.py): there PMT_ENV_INT("NAME", default) reads the environment ONCE per process and per call
// site (getenv is a linear scan of environ; the launch path of a 10-us kernel cannot afford several per call) and
// PMT_DBG(args, bit) tests an ablation bit of PMT_TC_DEBUG.  In the default (release) build both are compile-time
// constants, so a stray environment variable can neither change a tile configuration nor make a kernel skip work.
int env_int_uncached(const char* name, int dflt);
#ifdef PMT_DEV_KNOBS
#define PMT_ENV_INT(name, dflt) ([]() -> int { static const int v = ::pmt::env_int_uncached(name, dflt); return v; }())
#define PMT_DBG(args, bit) ((((args).debug) & (bit)) != 0)
#else
#define PMT_ENV_INT(name, dflt) (dflt)
#define PMT_DBG(args, bit) (false)
#endif

// 4-D fp32 tensor map over a dense (B,C,H,W) tensor, box = (box_w, 1, box_c, 1), zero OOB fill.
// Returns 0 on success (error text set otherwise).
int make_tmap_nchw(CUtensorMap* map, const float* base, int B, int C, int H, int W, int box_w,
                   int box_c);

// ---- device-side PTX wrappers ---------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Same, but a waiter that is not on the critical path sleeps between polls so that it does not compete with the
// working warps for issue slots and for the shared-memory pipeline that serves mbarrier operations.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns = 64) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// TMA: 4-D tiled load global -> shared, completion on an mbarrier (SASS: UTMALDG).
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                            int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
// TMA prefetch of a 4-D box into L2 only (no shared memory, no barrier): hides DRAM latency for a later tma_load_4d.
__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// cp.async 16-byte copy with zero fill when !valid (src-size 0).  (SASS: LDGSTS)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(sz)
               : "memory");
}
// 4-byte variant for rows whose start is not 16-byte aligned
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// streaming (evict-first) stores for write-once outputs
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
#endif  // __CUDACC__

}  // namespace pmt
