// pv_common.cuh — shared device/host helpers for the sm_100a kernels.
//
// Everything here is written for Blackwell (sm_100a) only: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (MMA / TMEM alloc / TMEM load) as inline PTX.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ---------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------
#define PV_OK 0
#define PV_ERR_INVALID (-1)
#define PV_ERR_CUDA (-2)
#define PV_ERR_UNSUPPORTED (-3)
#define PV_ERR_DEVICE_TIMEOUT (-4)

void pv_set_error(const char* fmt, ...);

#define PV_CUDA_CHECK(expr)                                                     \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      pv_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                   cudaGetErrorString(_e));                                     \
      return PV_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)

#define PV_REQUIRE(cond, ...)                                                   \
  do {                                                                          \
    if (!(cond)) {                                                              \
      pv_set_error(__VA_ARGS__);                                                \
      return PV_ERR_INVALID;                                                    \
    }                                                                           \
  } while (0)

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t pv_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint4 pv_lds128(uint32_t smem_addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
  return v;
}

// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool pv_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void pv_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pv_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void pv_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pv_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pv_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pv_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void pv_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pv_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool pv_mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(pv_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU (a hung box is a lost
// lease).  On timeout we raise a flag in global memory and trap.
__device__ __forceinline__ void pv_mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag,
                                             int code) {
  if (pv_mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!pv_mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}

// Same, for roles that are not on the critical path (deep rings ahead of them): sleep between polls so
// the spin loop does not compete for issue slots with the warps doing the work.
__device__ __forceinline__ void pv_mbar_wait_backoff(uint64_t* bar, uint32_t parity, int* err_flag, int code,
                                                     unsigned ns) {
  if (pv_mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  unsigned polls = 0;
  while (!pv_mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if ((++polls & 63u) == 0u && clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}

// ---- TMA ---------------------------------------------------------------------
__device__ __forceinline__ void pv_tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void pv_tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar,
                                               int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(pv_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(pv_smem_u32(bar)),
        "r"(c0), "r"(c1)
      : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------
__device__ __forceinline__ void pv_tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void pv_tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp
__device__ __forceinline__ void pv_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   pv_smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pv_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// single thread: D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void pv_umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate));
}
// warp-uniform variant: every lane executes the instruction slot, only lanes with `issue` != 0 issue the
// MMA (no divergent branch around it, so operands can stay in uniform registers)
__device__ __forceinline__ void pv_umma_bf16_pred(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(issue));
}
__device__ __forceinline__ void pv_umma_commit_pred(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(pv_smem_u32(bar)),
      "r"(issue)
      : "memory");
}
// single thread: arrive on mbarrier when all previously issued MMAs complete
__device__ __forceinline__ void pv_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   pv_smem_u32(bar))
               : "memory");
}
// whole warp: 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void pv_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// whole warp: write zeros to 32 lanes x 16 consecutive 32-bit columns (accumulator rows that are handed back cleared, so
// that every MMA can accumulate)
__device__ __forceinline__ void pv_tmem_st16_zero(uint32_t taddr) {
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void pv_tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pv_tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, 8-row groups `sbo`
// bytes apart (cute::UMMA::SmemDescriptor, sm_100: version field = 1).
//   layout_type: 0 none, 2 = 128B, 4 = 64B, 6 = 32B swizzle.
__device__ __forceinline__ uint64_t pv_umma_desc(uint32_t smem_addr, uint32_t sbo_bytes,
                                                 uint32_t layout_type, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // [0,14)
  d |= static_cast<uint64_t>(1) << 16;                            // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;    // [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // version = 1 (Blackwell)
  d |= static_cast<uint64_t>(base_offset & 7) << 49;              // [49,52)
  d |= static_cast<uint64_t>(layout_type & 7) << 61;              // [61,64)
  return d;
}

// one 256-bit global store (STG.E.256, sm_100+): a lane's 32 bytes leave as ONE full sector instead of two half-sector
// 128-bit stores (the epilogues write 32 / 64 / 96 bytes per pixel with lanes on neighbouring pixels)
__device__ __forceinline__ void pv_stg256(void* gptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                          uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(gptr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}

__device__ __forceinline__ uint32_t pv_pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

#endif  // __CUDACC__

// cudaFuncSetAttribute is per DEVICE: `flags` is a per-kernel static bitmask, bit d = attribute set on device d
static inline bool pv_attr_needed(unsigned long long* flags) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (*flags & bit) return false;
  *flags |= bit;
  return true;
}
