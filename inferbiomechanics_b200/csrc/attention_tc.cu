// Attention backward on the 5th-generation tensor cores (tcgen05) for short windows: T <= 64, head_dim 64.
//
// Replaces the autograd backward of nn.MultiheadAttention's scaled-dot-product core
// (/root/reference/src/models/TransformerBaseline.py:12-13,29) for the denoiser's 8 x 64 heads.
//
// One CTA iteration handles TWO windows of one head as a single 128-row problem with a block-diagonal mask:
// rows 0-63 are window a (T valid, zero pad rows), rows 64-127 window b.  All five products run as
// M = 128 tcgen05.mma with operands read straight from shared memory and accumulators in TMEM —
//     S  = Q K^T   (128 x 128 x 64)      dP = dO V^T  (128 x 128 x 64)
//     dV = P^T dO  (128 x  64 x 128)     dK = dS^T Q  (128 x  64 x 128)     dQ = dS K (128 x 64 x 128)
// — the cross-window blocks of S / dP are never read, P and dS are written block-diagonal (zero cross blocks),
// so the stacked products are exactly the two per-window products.  TMEM lane == row: the softmax, the row sums
// D = sum(P o dP) and dS are computed by the four threads that share a lane (16 keys each, three floats exchanged
// through shared memory, no shuffles).  Every smem tile is a
// SWIZZLE_128B [rows][128 B] tile that serves as K-major operand in one product and MN-major operand in
// another (e.g. dO is A of dP and B of dV) without any transposition.
//
// The legacy-tensor-core kernel (attention.cu: ldmatrix + mma.sync) was bound by the LSU data pipe (73 % busy
// with the operand fragments of four warps that each re-read K, V, Q, dO; profiles/r01b) at 44 % of the HBM
// roofline; here the LSU only carries the softmax outputs and the result staging.
#include <cuda.h>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

namespace ibm {
namespace attn_tc {

using namespace ptx;

constexpr int kThreads = 512;             // 16 warps: TMEM lane quadrant (warp & 3) x column quarter (warp >> 2)
constexpr int HD = 64;
constexpr int kTile = 128 * 128;          // one [128 rows][128 B] tile = 16 KB
// shared memory: 2 stages x {Q, K, V, dO} (also the staging of dQ, dK, dV) + P, dS (two 64-key chunks each) + row exchange
constexpr int kSmem = 1024 + 2 * 4 * kTile + 2 * 2 * kTile + 3 * 128 * 4 * 4 + 256;
// TMEM columns: S 0..127, dP 128..255, dQ 256..319, dV 320..383, dK 384..447
constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256, kColDV = 320, kColDK = 384;

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* r) { tmem_ld_32x16(taddr, reinterpret_cast<uint32_t*>(r)); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void add_bf16x8(float (&a)[8], uint4 u) {
  a[0] += __uint_as_float(u.x << 16); a[1] += __uint_as_float(u.x & 0xffff0000u);
  a[2] += __uint_as_float(u.y << 16); a[3] += __uint_as_float(u.y & 0xffff0000u);
  a[4] += __uint_as_float(u.z << 16); a[5] += __uint_as_float(u.z & 0xffff0000u);
  a[6] += __uint_as_float(u.w << 16); a[7] += __uint_as_float(u.w & 0xffff0000u);
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQKV, int T, int H, int64_t n_win, int64_t kv_off, float scale,
                   float* __restrict__ dbias) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage st: Q | K | V | dO at smem + st * 4 * kTile
  uint8_t* Ps = smem + 8 * kTile;                     // [2 key chunks][128 rows][128 B]
  uint8_t* dSs = Ps + 2 * kTile;
  float* xch = reinterpret_cast<float*>(dSs + 2 * kTile);               // [3][128 rows][4 quarters]: row max, sum, sum(e*dP)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(xch + 3 * 128 * 4);  // [2]
  uint64_t* s_bar = full_bar + 2;                     // S, dP complete
  uint64_t* o_bar = s_bar + 1;                        // dQ, dK, dV complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % H;
  const int64_t cta = blockIdx.x / H, ctas = gridDim.x / H;
  const int64_t n_pairs = (n_win + 1) >> 1;

  // zero everything once: pad rows (>= T of each 64-row half) of the operand stages stay zero for the whole kernel, and
  // so do the cross-window blocks of P and dS (only the diagonal blocks are ever written)
  for (int i = tid; i < 12 * kTile / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQKV);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();                           // the zero fill is visible to the TMA / MMA (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int q = warp & 3, cq4 = warp >> 2;            // TMEM lane quadrant, column quarter (16 keys / 16 head columns)
  const int r = q * 32 + lane;                        // this thread's row == TMEM lane (shared by 4 threads)
  const int ri = r & 63;                              // row inside its window
  const int own = r >> 6;                             // which window of the pair / which 64-key chunk is "ours"
  const int rsw = r & 7;
  const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
  const float scale_log2 = scale * 1.4426950408889634f;
  const uint32_t tx_bytes = 8u * (uint32_t)T * 128u;
  const uint32_t xrow = smem_u32(xch) + r * 16;       // this row's 4 exchange slots (16 B), three planes 2 KB apart
  // bias-gradient partial sums: 16-byte piece (tid & 7) of rows (tid >> 3) and 64 + (tid >> 3) of the staged results
  float cq[8], ck[8], cv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cq[j] = ck[j] = cv[j] = 0.f;

  auto issue_loads = [&](int64_t pair, int st) {      // thread 0
    uint8_t* base = smem + st * 4 * kTile;
    mbar_arrive_expect_tx(&full_bar[st], tx_bytes);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int64_t w = 2 * pair + it;
      if (w >= n_win) w = n_win - 1;                  // odd tail: a valid window again, its results are not stored
      const int32_t row0 = (int32_t)(w * T);
      uint8_t* dst = base + it * 64 * 128;
      tma_load_2d(dst, &tmQKV, &full_bar[st], h * HD, row0);
      tma_load_2d(dst + kTile, &tmQKV, &full_bar[st], (int32_t)kv_off + h * HD, row0);
      tma_load_2d(dst + 2 * kTile, &tmQKV, &full_bar[st], (int32_t)(2 * kv_off) + h * HD, row0);
      tma_load_2d(dst + 3 * kTile, &tmDO, &full_bar[st], h * HD, row0);
    }
  };
  // Shared-memory descriptors are built once: a k step only advances the 14-bit start-address field (by 32 B >> 4 = 2
  // for K-major operands, 2048 B >> 4 = 128 for MN-major ones), so the single issuing thread spends one add per operand
  // per MMA instead of rebuilding descriptors — with 32-clock MMAs the issue rate is what paces the tensor pipe.
  const uint64_t dK_Q0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);                  // stage 0 Q as K-major operand
  const uint64_t dM_Q0 = make_smem_desc_sw128(smem_u32(smem), kTile, 1024);              // stage 0 Q as MN-major operand
  constexpr uint64_t kTileStep = kTile >> 4, kStageStep = (4 * kTile) >> 4;
  const uint64_t dM_P = make_smem_desc_sw128(smem_u32(Ps), kTile, 1024);
  const uint64_t dM_dS = make_smem_desc_sw128(smem_u32(dSs), kTile, 1024);
  const uint64_t dK_dS = make_smem_desc_sw128(smem_u32(dSs), 0, 1024);
  const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
  // S = Q K^T, dP = dO V^T of the pair in stage st: K-major operands, K = 64 = 4 steps of 16        (thread 0)
  auto issue_scores = [&](int st, int use) {
    mbar_wait(&full_bar[st], (uint32_t)(use & 1));
    tc_fence_after();
    const uint64_t q = dK_Q0 + st * kStageStep;
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_bf16(tmem_base + kColS, q + 2 * k, q + kTileStep + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_bf16(tmem_base + kColDP, q + 3 * kTileStep + 2 * k, q + 2 * kTileStep + 2 * k, idesc_s, k > 0 ? 1u : 0u);
    umma_commit(s_bar);
  };

  if (tid == 0 && cta < n_pairs) {
    issue_loads(cta, 0);
    issue_scores(0, 0);
  }
  uint32_t ph_s = 0, ph_o = 0;
  int it_n = 0;
  for (int64_t pair = cta; pair < n_pairs; pair += ctas, ++it_n) {
    const int st = it_n & 1;
    uint8_t* Qs = smem + st * 4 * kTile;
    uint8_t* Ks = Qs + kTile;
    uint8_t* Vs = Ks + kTile;
    uint8_t* dOs = Vs + kTile;
    const bool second_valid = 2 * pair + 1 < n_win;
    const bool has_next = pair + ctas < n_pairs;

    if (tid == 0 && has_next) {
      // the other stage was the staging area of the previous iteration's result stores: they must have read it
      tma_wait_group_read<0>();
      issue_loads(pair + ctas, st ^ 1);
    }

    // ---- softmax / dS: four threads per row, 16 keys each ----
    mbar_wait(s_bar, ph_s);
    ph_s ^= 1u;
    tc_fence_after();
    float s[16], dp[16];
    tmem_ld_x16(lane_base + kColS + own * 64 + cq4 * 16, s);
    tmem_ld_x16(lane_base + kColDP + own * 64 + cq4 * 16, dp);
    tmem_ld_wait();
    tc_fence_before();
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (cq4 * 16 + j >= T) s[j] = -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
    sts_f(xrow + cq4 * 4, mx);
    named_bar_sync(1 + q, 128);                        // the four warps of this lane quadrant
    {
      const float4 m4 = lds_f4(xrow);                  // keys 0..15 always hold a valid key: the maximum is finite
      mx = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));
    }
    const float off = mx * scale_log2;
    float l = 0.f, ed = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      s[j] = ex2(fmaf(s[j], scale_log2, -off));        // masked keys: ex2(-inf) = 0
      l += s[j];
      ed = fmaf(s[j], dp[j], ed);
    }
    sts_f(xrow + 2048 + cq4 * 4, l);
    sts_f(xrow + 4096 + cq4 * 4, ed);
    named_bar_sync(1 + q, 128);
    {
      const float4 l4 = lds_f4(xrow + 2048), e4 = lds_f4(xrow + 4096);
      l = (l4.x + l4.y) + (l4.z + l4.w);
      ed = (e4.x + e4.y) + (e4.z + e4.w);
    }
    const float inv = ri < T ? 1.f / l : 0.f;          // padded query rows contribute nothing
    const float dsc = ed * inv;                        // D = sum_j P_j dP_j
    {
      const uint32_t prow = smem_u32(Ps) + own * kTile + r * 128, srow = smem_u32(dSs) + own * kTile + r * 128;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        uint32_t pw[4], sw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 8 * jj + 2 * e;
          const float p0 = s[j] * inv, p1 = s[j + 1] * inv;
          pw[e] = pack_bf16x2(p0, p1);
          sw[e] = pack_bf16x2(p0 * (dp[j] - dsc) * scale, p1 * (dp[j + 1] - dsc) * scale);
        }
        const uint32_t po = (uint32_t)(((cq4 * 2 + jj) ^ rsw) << 4);
        sts16(prow + po, make_uint4(pw[0], pw[1], pw[2], pw[3]));
        sts16(srow + po, make_uint4(sw[0], sw[1], sw[2], sw[3]));
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      // ---- dV = P^T dO, dK = dS^T Q : A MN-major (M = keys: two 64-wide atoms 16 KB apart), B MN-major, K = 128 queries ----
      const uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t qm = dM_Q0 + st * kStageStep;                     // this stage's Q | K | V | dO as MN-major operands
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem_base + kColDV, dM_P + 128 * k, qm + 3 * kTileStep + 128 * k, idesc_t, k > 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem_base + kColDK, dM_dS + 128 * k, qm + 128 * k, idesc_t, k > 0 ? 1u : 0u);
      // ---- dQ = dS K : A K-major (K = 128 keys = two chunks), B MN-major ----
      const uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_bf16(tmem_base + kColDQ, dK_dS + (k >> 2) * kTileStep + (k & 3) * 2, qm + kTileStep + 128 * k, idesc_q, k > 0 ? 1u : 0u);
      umma_commit(o_bar);
      // S / dP are free (every thread passed the barrier above): start the next pair's scores behind these products,
      // so they are ready when this pair's results have been drained
      if (has_next) issue_scores(st ^ 1, (it_n + 1) >> 1);
    }

    // ---- results: TMEM -> bf16 -> the (now dead) Q / K / V tiles -> TMA stores (one wait for all three products: waiting
    //      per product so that the drain overlaps the remaining MMAs measured slower, 470 vs 439 us) ----
    uint8_t* Gq = Qs;
    uint8_t* Gk = Ks;
    uint8_t* Gv = Vs;
    mbar_wait(o_bar, ph_o);
    ph_o ^= 1u;
    tc_fence_after();
    {
      float vq[16], vk[16], vv[16];
      tmem_ld_x16(lane_base + kColDQ + cq4 * 16, vq);
      tmem_ld_x16(lane_base + kColDK + cq4 * 16, vk);
      tmem_ld_x16(lane_base + kColDV + cq4 * 16, vv);
      tmem_ld_wait();
      if (ri < T) {
        const uint32_t ro = r * 128;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const uint32_t po = ro + (uint32_t)(((cq4 * 2 + jj) ^ rsw) << 4);
          sts16(smem_u32(Gq) + po, make_uint4(pack_bf16x2(vq[8 * jj], vq[8 * jj + 1]), pack_bf16x2(vq[8 * jj + 2], vq[8 * jj + 3]),
                                              pack_bf16x2(vq[8 * jj + 4], vq[8 * jj + 5]), pack_bf16x2(vq[8 * jj + 6], vq[8 * jj + 7])));
          sts16(smem_u32(Gk) + po, make_uint4(pack_bf16x2(vk[8 * jj], vk[8 * jj + 1]), pack_bf16x2(vk[8 * jj + 2], vk[8 * jj + 3]),
                                              pack_bf16x2(vk[8 * jj + 4], vk[8 * jj + 5]), pack_bf16x2(vk[8 * jj + 6], vk[8 * jj + 7])));
          sts16(smem_u32(Gv) + po, make_uint4(pack_bf16x2(vv[8 * jj], vv[8 * jj + 1]), pack_bf16x2(vv[8 * jj + 2], vv[8 * jj + 3]),
                                              pack_bf16x2(vv[8 * jj + 4], vv[8 * jj + 5]), pack_bf16x2(vv[8 * jj + 6], vv[8 * jj + 7])));
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        if (it == 1 && !second_valid) break;
        const int32_t row0 = (int32_t)((2 * pair + it) * T);
        tma_store_2d(&tmDQKV, Gq + it * 64 * 128, h * HD, row0);
        tma_store_2d(&tmDQKV, Gk + it * 64 * 128, (int32_t)kv_off + h * HD, row0);
        tma_store_2d(&tmDQKV, Gv + it * 64 * 128, (int32_t)(2 * kv_off) + h * HD, row0);
      }
      tma_commit_group();
    }
    if (dbias != nullptr) {
      // in_proj bias gradient from the staged (rounded) rows: this thread owns 16-byte piece (tid & 7) of rows g and 64 + g
      const int j = tid & 7, g = tid >> 3;
      if (g < T) {
        const uint32_t o0 = (uint32_t)(g * 128 + ((j ^ (g & 7)) << 4));
        add_bf16x8(cq, lds16(smem_u32(Gq) + o0));
        add_bf16x8(ck, lds16(smem_u32(Gk) + o0));
        add_bf16x8(cv, lds16(smem_u32(Gv) + o0));
        if (second_valid) {
          add_bf16x8(cq, lds16(smem_u32(Gq) + o0 + 64 * 128));
          add_bf16x8(ck, lds16(smem_u32(Gk) + o0 + 64 * 128));
          add_bf16x8(cv, lds16(smem_u32(Gv) + o0 + 64 * 128));
        }
      }
    }
    __syncthreads();                                  // staging reads done before this stage is loaded again
  }
  if (tid == 0) tma_wait_group<0>();
  if (dbias != nullptr) {
    // lanes that share (tid & 7) hold partial sums of the same 8 columns: fold the warp's 4 row groups, then one atomic
    // per column per warp
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 8);  cq[j] += __shfl_xor_sync(0xffffffffu, cq[j], 16);
      ck[j] += __shfl_xor_sync(0xffffffffu, ck[j], 8);  ck[j] += __shfl_xor_sync(0xffffffffu, ck[j], 16);
      cv[j] += __shfl_xor_sync(0xffffffffu, cv[j], 8);  cv[j] += __shfl_xor_sync(0xffffffffu, cv[j], 16);
    }
    if (lane < 8) {
      const int c = h * HD + lane * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(dbias + c + j, cq[j]);
        atomicAdd(dbias + kv_off + c + j, ck[j]);
        atomicAdd(dbias + 2 * kv_off + c + j, cv[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace attn_tc

// returns IBM_E_UNSUPPORTED when the shape is outside this kernel's domain (the caller falls back to mma.sync)
int attention_bwd_tc(const void* qkv, int64_t ld, int64_t kv_off, const void* d_o, int64_t ldo, void* dqkv, int64_t n_win, int T, int H,
                     int head_dim, float scale, float* dbias, cudaStream_t s) {
  using namespace attn_tc;
  if (head_dim != HD || T > 64 || T < 1 || scale <= 0.f) return IBM_E_UNSUPPORTED;
  if (n_win * T >= (1ll << 31) || 3 * kv_off >= (1ll << 31)) return IBM_E_UNSUPPORTED;
  const int64_t rows = n_win * T;
  const int64_t cols = 2 * kv_off + (int64_t)H * HD;
  CUtensorMap mq, md, mg;
  int rc = make_map(&mq, qkv, false, cols, rows, ld, 64, (uint32_t)T);
  if (rc) return rc;
  rc = make_map(&md, d_o, false, (int64_t)H * HD, rows, ldo, 64, (uint32_t)T);
  if (rc) return rc;
  rc = make_map(&mg, dqkv, false, cols, rows, ld, 64, (uint32_t)T);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    IBM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int64_t n_pairs = (n_win + 1) / 2;
  int64_t grid = (int64_t)sm_count() / H * H;
  if (grid < H) grid = H;
  if (grid > n_pairs * H) grid = n_pairs * H;
  attn_bwd_tc_kernel<<<(unsigned)grid, kThreads, kSmem, s>>>(mq, md, mg, T, H, n_win, kv_off, scale, dbias);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

}  // namespace ibm
