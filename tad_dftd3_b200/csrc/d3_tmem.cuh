// Tensor Memory (TMEM, 256 kB per SM on sm_100a) as a software-managed per-lane scratch.
//
// The fused per-structure kernel has no tensor-core work, so the SM's tensor memory would sit
// idle.  It is used here as a cache between two sweeps over the same pairs: the CN sweep
// stores one value per (lane, partner step) -- the derivative factor of the counting function
// -- and the CN-backward sweep reads it back instead of recomputing rsqrt + exp + reciprocal
// (see d3_molecule_v2.cuh).  Shared memory is full (4 CTAs x 55 kB) and an L2-resident global
// scratch costs more latency than the recomputation (profiles/r1f_summary.md); TMEM is
// lane-private, which is exactly the access pattern: a lane re-reads what it wrote itself.
//
// Layout: 512 columns x 128 lanes x 32 bit per SM.  A CTA of four warps allocates kCols columns;
// warp w owns lanes 32 (w % 4) .. +31 (the hardware restricts a warp to that quarter), thread
// `lane` of the warp reads and writes TMEM lane 32 (w % 4) + lane (shape .32x32b), columns are
// consecutive 32-bit words: a double takes two columns.  All instructions are warp-collective
// (.sync.aligned): callers keep the surrounding control flow warp-uniform.
#pragma once
#include <stdint.h>

namespace d3 {
namespace tmem {

constexpr int kCols = 128;   // columns per CTA: 4 resident CTAs per SM share the 512 columns

__device__ __forceinline__ void alloc(uint32_t* smem_slot) {   // one full warp; writes the base address to *smem_slot
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_slot);
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"((uint32_t)kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t base) {        // one full warp, after every warp is done with TMEM
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"((uint32_t)kCols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// address of column `col` in the lane quarter of warp `wid`
__device__ __forceinline__ uint32_t addr(uint32_t base, int wid, int col) {
  return base + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)col;
}

__device__ __forceinline__ void st_words(uint32_t a, const uint32_t (&r)[2]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void st_words(uint32_t a, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void st_words(uint32_t a, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void ld_words(uint32_t a, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_words(uint32_t a, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a)
               : "memory");
}
__device__ __forceinline__ void ld_words(uint32_t a, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(a)
               : "memory");
}

// NV values of type float / double per lane <-> NV * sizeof(S) / 4 consecutive columns
template <int NV> __device__ __forceinline__ void store(uint32_t a, const double (&v)[NV]) {
  uint32_t r[2 * NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) { r[2 * k] = (uint32_t)__double2loint(v[k]); r[2 * k + 1] = (uint32_t)__double2hiint(v[k]); }
  st_words(a, r);
}
template <int NV> __device__ __forceinline__ void load(uint32_t a, double (&v)[NV]) {
  uint32_t r[2 * NV];
  ld_words(a, r);
  wait_ld();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
template <int NV> __device__ __forceinline__ void store(uint32_t a, const float (&v)[NV]) {
  uint32_t r[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) r[k] = __float_as_uint(v[k]);
  st_words(a, r);
}
template <int NV> __device__ __forceinline__ void load(uint32_t a, float (&v)[NV]) {
  uint32_t r[NV];
  ld_words(a, r);
  wait_ld();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = __uint_as_float(r[k]);
}

}  // namespace tmem
}  // namespace d3
