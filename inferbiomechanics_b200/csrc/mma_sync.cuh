// Warp-level mma.sync / ldmatrix / cp.async helpers shared by the attention kernels (attention.cu, attention_bwd_long.cu).
#pragma once

#include "common.cuh"

namespace ibm {
namespace attn {

constexpr int kPad = 8;             // bf16 elements of row padding → conflict-free fragment loads

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// B fragment (16 k x 8 n) from a row-major [k][n] tile: lanes 0-15 pass &X[k0 + lane][n0]
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
// A fragment (16 m x 16 k) of Y^T from a row-major Y[k][m] tile
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// four 8x8 b16 matrices, row-major reads: lane l passes the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two 8x8 matrices, row-major: lanes 0-15 pass the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// acc[16 x HD] += A[16 x 16] · B[16 x HD] for one 16-row block of B starting at `Brows` (row-major, stride LB)
template <int HD, int LB>
__device__ __forceinline__ void mma_ab16(float (&acc)[HD / 8][4], const uint32_t (&a)[4], const __nv_bfloat16* Brows, int lane) {
  const __nv_bfloat16* bp = Brows + (lane & 15) * LB + (lane >> 4) * 8;
#pragma unroll
  for (int j2 = 0; j2 < HD / 16; ++j2) {
    uint32_t b[4];
    ldsm_x4_trans(b, bp + j2 * 16);
    mma16816(acc[2 * j2], a, b[0], b[1]);
    mma16816(acc[2 * j2 + 1], a, b[2], b[3]);
  }
}
}  // namespace attn
}  // namespace ibm
