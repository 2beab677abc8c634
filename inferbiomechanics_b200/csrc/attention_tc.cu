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

constexpr int kComputeThreads = 512;      // 16 warps: TMEM lane quadrant (warp & 3) x column quarter (warp >> 2)
// + five control warps, one issuing thread each.  A tcgen05.mma costs its issuing thread ~65 clocks and a TMA operation
// ~130 (clock64 trace of one iteration: 24 MMAs 1 620 clocks, 6 loads 780), far more than these small MMAs keep the tensor
// pipe busy (32-64 clocks), so one control thread paced the whole kernel at ~7 000 clocks per pair; the issue work is
// spread over: loader | dV + S | dK + dP | dQ + storer (20 warps in all: a 21st would drop the register budget of the
// compute warps from 96 to 80).
constexpr int kControlWarps = 4;
constexpr int kThreads = kComputeThreads + 32 * kControlWarps;
constexpr int HD = 64;
constexpr int kTile = 128 * 128;          // one [128 rows][128 B] tile = 16 KB
// P and dS are block-diagonal [128 queries][2 x 64 keys]: chunk 0 = [P_a ; 0], chunk 1 = [0 ; P_b].  The two chunks are
// laid 8 KB apart instead of 16, so the zero half of chunk 0 (rows 64-127) IS the zero half of chunk 1 (rows 0-63):
// [P_a 8 KB | zeros 8 KB | P_b 8 KB] = 24 KB per matrix, and the descriptors simply use an 8 KB atom / chunk stride.
constexpr int kPD = 3 * 8192;
constexpr int kChunk = 8192;
// shared memory: 2 stages x {Q, K, dO} + one V tile + P, dS + result staging dQ | dK | dV + row exchange
constexpr int kOnes = 4096;                // a [16][128] tile of bf16 ones: B operand of the bias-gradient products
constexpr int kSmem = 1024 + 2 * 3 * kTile + kTile + 2 * kPD + 3 * kTile + kOnes + 3 * 128 * 4 * 4 + 256;
// TMEM columns: S 0..127, dP 128..255, dQ 256..319, dV 320..383, dK 384..447
constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256, kColDV = 320, kColDK = 384;
// bias gradient: column sums of the staged dQ | dK (lanes 0-63 | 64-127) at 448 and of dK | dV at 464, accumulated over
// every pair of the CTA by M = 128, N = 16 products  [dQ | dK]^T · ones,  [dK | dV]^T · ones  (two adjacent staging tiles
// form one MN-major operand) — no work for the compute warps, no registers
constexpr uint32_t kColBqk = 448, kColBkv = 464;

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

// Software pipeline over the window pairs of one head (one CTA per SM: the TMEM map below needs 448 of its 512 columns).
// The tensor pipe and the 16 compute warps work on different pairs at the same time:
//     tensor pipe :  ... | products(i-1) = dV,dK,dQ | scores(i) = S,dP | products(i) | scores(i+1) | ...
//     compute     :  ... | softmax(i) in registers  | fetch results(i-1), write P/dS(i) | stage + store results(i-1) | ...
// which works because the two groups of MMAs use disjoint TMEM columns, P/dS of pair i stay in registers until the
// products of pair i-1 have released the shared-memory P/dS tiles, and results are copied TMEM -> registers before the next
// products overwrite them.  Five control warps (one thread each) issue the TMA loads, the MMAs and the TMA stores.
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQKV, int T, int H, int64_t n_win, int64_t kv_off, float scale,
                   float* __restrict__ dbias, int rev) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage st: Q | K | dO at smem + st * 3 * kTile
  uint8_t* Vs = smem + 6 * kTile;                     // single: V is dead as soon as the scores of its pair are complete
  uint8_t* Ps = Vs + kTile;                           // [P_a | shared zero block | P_b], see kPD
  uint8_t* dSs = Ps + kPD;
  uint8_t* Gq = dSs + kPD;                            // result staging tiles
  uint8_t* Gk = Gq + kTile;
  uint8_t* Gv = Gk + kTile;
  uint8_t* Ones = Gv + kTile;                         // kOnes bytes of bf16 1.0
  float* xch = reinterpret_cast<float*>(Ones + kOnes);                  // [3][128 rows][4 quarters]: row max, sum, sum(e*dP)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(xch + 3 * 128 * 4);  // [2] Q, K, dO of a stage landed
  uint64_t* v_bar = full_bar + 2;                     // V landed
  uint64_t* s_bar = v_bar + 1;                        // S, dP complete
  uint64_t* o_bar = s_bar + 1;                        // dQ, dK, dV complete
  uint64_t* pd_bar = o_bar + 1;                       // P/dS of the pair are in shared memory, its predecessor's results in registers
  uint64_t* stg_bar = pd_bar + 1;                     // results staged
  uint64_t* sf_bar = stg_bar + 1;                     // staging tiles free again (TMA stores and bias products have read them)
  uint64_t* c_bar = sf_bar + 1;                       // all bias-gradient products complete
  uint64_t* sc_bar = c_bar + 1;                       // S / dP of the pair have been copied to registers: the columns are free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sc_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % H;
  const int64_t cta = blockIdx.x / H, ctas = gridDim.x / H;
  const int64_t n_pairs = (n_win + 1) >> 1;
  const int n_it = cta < n_pairs ? (int)((n_pairs - cta + ctas - 1) / ctas) : 0;      // pairs of this CTA
  // i-th pair of this CTA; rev: pairs are taken from the last one down (ibm_set_walk_order)
  const int64_t pair0 = rev ? n_pairs - 1 - cta : cta, pair_step = rev ? -ctas : ctas;
  auto pair_of = [&](int i) -> int64_t { return pair0 + (int64_t)i * pair_step; };

  // zero everything once: pad rows (>= T of each 64-row half) of the operand tiles stay zero for the whole kernel, and
  // so do the cross-window blocks of P and dS (only the diagonal blocks are ever written)
  for (int i = tid; i < (10 * kTile + 2 * kPD) / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < kOnes / 16; i += kThreads) reinterpret_cast<uint4*>(Ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (tid == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQKV);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    mbar_init(v_bar, 1);
    mbar_init(s_bar, 2);                              // S and dP are committed by two different warps
    mbar_init(o_bar, 3);                              // dV, dK, dQ by three
    mbar_init(pd_bar, kComputeThreads / 32);
    mbar_init(stg_bar, kComputeThreads / 32);
    mbar_init(sf_bar, dbias != nullptr ? 3 : 1);      // storer + the two warps that issue the bias products
    mbar_init(c_bar, 2);
    mbar_init(sc_bar, kComputeThreads / 32);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  fence_proxy_async_smem();                           // the zero fill is visible to the TMA / MMA (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kComputeThreads / 32) {
    // ===================================== control warps =====================================
    const int role = warp - kComputeThreads / 32;     // 0 loader, 1 dV + S, 2 dK + dP, 3 dQ + storer
    if (lane == 0 && n_it > 0) {
      const uint32_t tx_qkd = 6u * (uint32_t)T * 128u, tx_v = 2u * (uint32_t)T * 128u;
      auto window_row = [&](int i, int it) {
        int64_t w = 2 * pair_of(i) + it;
        if (w >= n_win) w = n_win - 1;                // odd tail: a valid window again, its results are not stored
        return (int32_t)(w * T);
      };
      // descriptors are built once; a k step only advances the 14-bit start-address field
      const uint64_t dK_Q0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);          // stage 0 Q as K-major operand
      const uint64_t dM_Q0 = make_smem_desc_sw128(smem_u32(smem), kTile, 1024);      // stage 0 Q as MN-major operand
      constexpr uint64_t kTileStep = kTile >> 4, kStageStep = (3 * kTile) >> 4, kChunkStep = kChunk >> 4;
      const uint64_t dK_V = make_smem_desc_sw128(smem_u32(Vs), 0, 1024);
      const uint64_t dM_P = make_smem_desc_sw128(smem_u32(Ps), kChunk, 1024);        // the two 64-key atoms are 8 KB apart
      const uint64_t dM_dS = make_smem_desc_sw128(smem_u32(dSs), kChunk, 1024);
      const uint64_t dK_dS = make_smem_desc_sw128(smem_u32(dSs), 0, 1024);
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1), idesc_q = make_idesc_bf16(128, 64, 0, 1);
      // bias gradient of pair i: D[col][*] += sum_rows G[row][col], A = two staged tiles as one MN-major operand (M = 2 x 64
      // columns, K = 128 rows), B = ones (K-major, N = 16); accumulates over all pairs; signals "staging free" when done
      const uint64_t dM_G = make_smem_desc_sw128(smem_u32(Gq), kTile, 1024);
      const uint64_t dK_1 = make_smem_desc_sw128(smem_u32(Ones), 0, 1024);
      const uint32_t idesc_b = make_idesc_bf16(128, 16, 1, 0);
      auto bias_products = [&](int i, uint32_t col, uint64_t a_desc) {
        mbar_wait(stg_bar, (uint32_t)(i & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_base + col, a_desc + 128 * k, dK_1 + (k >> 2) * 128 + (k & 3) * 2, idesc_b, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(sf_bar);
      };
      auto wait_inputs = [&](int i) {                 // Q, K, dO and V of pair i have landed
        mbar_wait(&full_bar[i & 1], (uint32_t)((i >> 1) & 1));
        mbar_wait(v_bar, (uint32_t)(i & 1));
        tc_fence_after();
      };

      if (role == 0) {
        // ---- loader ----
        auto load_qkd = [&](int i) {                  // Q, K, dO of pair i into stage i & 1
          uint8_t* base = smem + (i & 1) * 3 * kTile;
          mbar_arrive_expect_tx(&full_bar[i & 1], tx_qkd);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int32_t row0 = window_row(i, it);
            uint8_t* dst = base + it * 64 * 128;
            tma_load_2d(dst, &tmQKV, &full_bar[i & 1], h * HD, row0);
            tma_load_2d(dst + kTile, &tmQKV, &full_bar[i & 1], (int32_t)kv_off + h * HD, row0);
            tma_load_2d(dst + 2 * kTile, &tmDO, &full_bar[i & 1], h * HD, row0);
          }
        };
        auto load_v = [&](int i) {
          mbar_arrive_expect_tx(v_bar, tx_v);
#pragma unroll
          for (int it = 0; it < 2; ++it)
            tma_load_2d(Vs + it * 64 * 128, &tmQKV, v_bar, (int32_t)(2 * kv_off) + h * HD, window_row(i, it));
        };
        // Shared memory holds the operands of ~1.5 pairs (51 KB each): with HBM latency near 2 us that is only ~25 GB/s per
        // SM in flight, and every schedule of this kernel measured the same ~3.2 TB/s.  Pairs further ahead are therefore
        // pulled into L2 (no shared memory needed), so the shared-memory loads see L2 latency.
        constexpr int kAhead = 4;
        auto prefetch_pair = [&](int i) {
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int32_t row0 = window_row(i, it);
            tma_prefetch_l2_2d(&tmQKV, h * HD, row0);
            tma_prefetch_l2_2d(&tmQKV, (int32_t)kv_off + h * HD, row0);
            tma_prefetch_l2_2d(&tmQKV, (int32_t)(2 * kv_off) + h * HD, row0);
            tma_prefetch_l2_2d(&tmDO, h * HD, row0);
          }
        };
        load_qkd(0);
        load_v(0);
        if (n_it > 1) load_qkd(1);
        for (int i = 1; i < kAhead && i < n_it; ++i) prefetch_pair(i);
        for (int i = 0; i < n_it; ++i) {
          if (i + kAhead < n_it) prefetch_pair(i + kAhead);
          if (i + 1 < n_it) {                         // V(i) is dead once the scores of pair i are complete
            mbar_wait(s_bar, (uint32_t)(i & 1));
            load_v(i + 1);
          }
          if (i + 2 < n_it) {                         // Q, K, dO of pair i die with its products
            mbar_wait(o_bar, (uint32_t)(i & 1));
            load_qkd(i + 2);
          }
        }
      } else if (role == 1) {
        // ---- dV(i) = P^T dO (A, B MN-major, K = 128 queries), then S(i+1) = Q K^T (K-major, K = 64) ----
        wait_inputs(0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + kColS, dK_Q0 + 2 * k, dK_Q0 + kTileStep + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_bar);
        for (int i = 0; i < n_it; ++i) {
          // the next pair's scores start as soon as this pair's S / dP have been read out of TMEM (early in the softmax):
          // they are what the compute warps wait for next
          if (i + 1 < n_it) {
            wait_inputs(i + 1);
            mbar_wait(sc_bar, (uint32_t)(i & 1));
            tc_fence_after();
            const uint64_t q = dK_Q0 + ((i + 1) & 1) * kStageStep;
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + kColS, q + 2 * k, q + kTileStep + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(s_bar);
          }
          mbar_wait(pd_bar, (uint32_t)(i & 1));       // P/dS(i) written; results(i-1) are in registers
          tc_fence_after();
          const uint64_t qm = dM_Q0 + (i & 1) * kStageStep;
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16(tmem_base + kColDV, dM_P + 128 * k, qm + 2 * kTileStep + 128 * k, idesc_t, k > 0 ? 1u : 0u);
          umma_commit(o_bar);
          if (dbias != nullptr && i > 0) bias_products(i - 1, kColBqk, dM_G);
        }
        if (dbias != nullptr) {
          bias_products(n_it - 1, kColBqk, dM_G);
          umma_commit(c_bar);
        }
      } else if (role == 2) {
        // ---- dK(i) = dS^T Q, then dP(i+1) = dO V^T ----
        wait_inputs(0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + kColDP, dK_Q0 + 2 * kTileStep + 2 * k, dK_V + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_bar);
        for (int i = 0; i < n_it; ++i) {
          if (i + 1 < n_it) {
            wait_inputs(i + 1);
            mbar_wait(sc_bar, (uint32_t)(i & 1));
            tc_fence_after();
            const uint64_t q = dK_Q0 + ((i + 1) & 1) * kStageStep;
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + kColDP, q + 2 * kTileStep + 2 * k, dK_V + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(s_bar);
          }
          mbar_wait(pd_bar, (uint32_t)(i & 1));
          tc_fence_after();
          const uint64_t qm = dM_Q0 + (i & 1) * kStageStep;
#pragma unroll
          for (int k = 0; k < 8; ++k) umma_bf16(tmem_base + kColDK, dM_dS + 128 * k, qm + 128 * k, idesc_t, k > 0 ? 1u : 0u);
          umma_commit(o_bar);
          if (dbias != nullptr && i > 0) bias_products(i - 1, kColBkv, dM_G + kTileStep);
        }
        if (dbias != nullptr) {
          bias_products(n_it - 1, kColBkv, dM_G + kTileStep);
          umma_commit(c_bar);
        }
      } else {
        // ---- dQ(i) = dS K (A K-major, K = 128 keys = two chunks; B MN-major), then the stores of pair i-1's staged results ----
        auto store_results = [&](int i) {
          mbar_wait(stg_bar, (uint32_t)(i & 1));
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int64_t w = 2 * pair_of(i) + it;
            if (w >= n_win) break;
            const int32_t row0 = (int32_t)(w * T);
            tma_store_2d(&tmDQKV, Gq + it * 64 * 128, h * HD, row0);
            tma_store_2d(&tmDQKV, Gk + it * 64 * 128, (int32_t)kv_off + h * HD, row0);
            tma_store_2d(&tmDQKV, Gv + it * 64 * 128, (int32_t)(2 * kv_off) + h * HD, row0);
          }
          tma_commit_group();
          tma_wait_group_read<0>();
          mbar_arrive(sf_bar);                        // staging tiles may be rewritten
        };
        for (int i = 0; i < n_it; ++i) {
          mbar_wait(pd_bar, (uint32_t)(i & 1));
          tc_fence_after();
          const uint64_t qm = dM_Q0 + (i & 1) * kStageStep;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + kColDQ, dK_dS + (k >> 2) * kChunkStep + (k & 3) * 2, qm + kTileStep + 128 * k, idesc_q, k > 0 ? 1u : 0u);
          umma_commit(o_bar);
          if (i > 0) store_results(i - 1);
        }
        store_results(n_it - 1);
        tma_wait_group<0>();
      }
    }
  } else {
    // ===================================== compute warps =====================================
    const int q = warp & 3, cq4 = warp >> 2;          // TMEM lane quadrant, column quarter (16 keys / 16 head columns)
    const int r = q * 32 + lane;                      // this thread's row == TMEM lane (shared by 4 threads)
    const int ri = r & 63;                            // row inside its window
    const int own = r >> 6;                           // which window of the pair / which 64-key chunk is "ours"
    const int rsw = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale_log2 = scale * 1.4426950408889634f;
    const uint32_t xrow = smem_u32(xch) + r * 16;     // this row's 4 exchange slots (16 B), three planes 2 KB apart
    const uint32_t prow = smem_u32(Ps) + own * kChunk + r * 128, srow = smem_u32(dSs) + own * kChunk + r * 128;
    const uint32_t po0 = (uint32_t)(((cq4 * 2) ^ rsw) << 4), po1 = (uint32_t)(((cq4 * 2 + 1) ^ rsw) << 4);

    // fetch pair i's results into registers is interleaved below; this finishes the job: pack, stage, hand to the control warp
    auto stage_results = [&](int i, const float (&vq)[16], const float (&vk)[16], const float (&vv)[16]) {
      if (i > 0) mbar_wait(sf_bar, (uint32_t)((i - 1) & 1));           // the previous pair's stores have read the staging tiles
      // odd tail: the second window of the last pair is a duplicate — its rows are staged as zeros (they are not stored,
      // and must not enter the bias gradient)
      const bool dup = own == 1 && 2 * pair_of(i) + 1 >= n_win;
      if (ri < T) {
        const uint32_t ro = r * 128;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const uint32_t po = ro + (jj == 0 ? po0 : po1);
          const uint4 z = make_uint4(0, 0, 0, 0);
          sts16(smem_u32(Gq) + po, dup ? z : make_uint4(pack_bf16x2(vq[8 * jj], vq[8 * jj + 1]), pack_bf16x2(vq[8 * jj + 2], vq[8 * jj + 3]),
                                                        pack_bf16x2(vq[8 * jj + 4], vq[8 * jj + 5]), pack_bf16x2(vq[8 * jj + 6], vq[8 * jj + 7])));
          sts16(smem_u32(Gk) + po, dup ? z : make_uint4(pack_bf16x2(vk[8 * jj], vk[8 * jj + 1]), pack_bf16x2(vk[8 * jj + 2], vk[8 * jj + 3]),
                                                        pack_bf16x2(vk[8 * jj + 4], vk[8 * jj + 5]), pack_bf16x2(vk[8 * jj + 6], vk[8 * jj + 7])));
          sts16(smem_u32(Gv) + po, dup ? z : make_uint4(pack_bf16x2(vv[8 * jj], vv[8 * jj + 1]), pack_bf16x2(vv[8 * jj + 2], vv[8 * jj + 3]),
                                                        pack_bf16x2(vv[8 * jj + 4], vv[8 * jj + 5]), pack_bf16x2(vv[8 * jj + 6], vv[8 * jj + 7])));
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(stg_bar);            // control: TMA stores and the bias-gradient products
    };

    for (int i = 0; i < n_it; ++i) {
      // ---- softmax / dS of pair i in registers: four threads per row, 16 keys each ----
      mbar_wait(s_bar, (uint32_t)(i & 1));
      tc_fence_after();
      float s[16], dp[16];
      tmem_ld_x16(lane_base + kColS + own * 64 + cq4 * 16, s);
      tmem_ld_x16(lane_base + kColDP + own * 64 + cq4 * 16, dp);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sc_bar);              // control: the next pair's scores may overwrite S / dP
      // One exchange round: every quarter publishes its local maximum m_q, sum l_q = sum exp(s - m_q) and
      // e_q = sum exp(s - m_q) dP, then rescales by exp(m_q - M) like an online softmax merge.
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (cq4 * 16 + j >= T) s[j] = -INFINITY;
        mx = fmaxf(mx, s[j]);
      }
      // a quarter that holds no valid key (T <= 16 * cq4) publishes m = -inf, l = e = 0 and scales by 0 below
      const float off = mx == -INFINITY ? 0.f : mx * scale_log2;
      float l = 0.f, ed = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s[j] = ex2(fmaf(s[j], scale_log2, -off));      // masked keys: ex2(-inf) = 0
        l += s[j];
        ed = fmaf(s[j], dp[j], ed);
      }
      sts_f(xrow + cq4 * 4, mx);
      sts_f(xrow + 2048 + cq4 * 4, l);
      sts_f(xrow + 4096 + cq4 * 4, ed);
      named_bar_sync(1 + q, 128);                      // the four warps of this lane quadrant
      float own_scale;
      {
        const float4 m4 = lds_f4(xrow), l4 = lds_f4(xrow + 2048), e4 = lds_f4(xrow + 4096);
        const float M = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w));      // quarter 0 always holds a valid key: finite
        const float Ml = M * scale_log2;
        const float f0 = ex2(fmaf(m4.x, scale_log2, -Ml)), f1 = ex2(fmaf(m4.y, scale_log2, -Ml));
        const float f2 = ex2(fmaf(m4.z, scale_log2, -Ml)), f3 = ex2(fmaf(m4.w, scale_log2, -Ml));
        l = fmaf(l4.x, f0, fmaf(l4.y, f1, fmaf(l4.z, f2, l4.w * f3)));
        ed = fmaf(e4.x, f0, fmaf(e4.y, f1, fmaf(e4.z, f2, e4.w * f3)));
        own_scale = cq4 == 0 ? f0 : (cq4 == 1 ? f1 : (cq4 == 2 ? f2 : f3));
      }
      const float inv = ri < T ? 1.f / l : 0.f;        // padded query rows contribute nothing
      const float dsc = ed * inv;                      // D = sum_j P_j dP_j
      const float pscale = inv * own_scale;            // exp(s - m_q) -> P
      uint32_t pw[8], sw[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float p0 = s[2 * e] * pscale, p1 = s[2 * e + 1] * pscale;
        pw[e] = pack_bf16x2(p0, p1);
        sw[e] = pack_bf16x2(p0 * (dp[2 * e] - dsc) * scale, p1 * (dp[2 * e + 1] - dsc) * scale);
      }
      // ---- the products of pair i-1 are complete: take its results out of TMEM, which also frees the P/dS tiles ----
      float vq[16], vk[16], vv[16];
      if (i > 0) {
        mbar_wait(o_bar, (uint32_t)((i - 1) & 1));
        tc_fence_after();
        tmem_ld_x16(lane_base + kColDQ + cq4 * 16, vq);
        tmem_ld_x16(lane_base + kColDK + cq4 * 16, vk);
        tmem_ld_x16(lane_base + kColDV + cq4 * 16, vv);
        tmem_ld_wait();
      }
      tc_fence_before();
      sts16(prow + po0, make_uint4(pw[0], pw[1], pw[2], pw[3]));
      sts16(prow + po1, make_uint4(pw[4], pw[5], pw[6], pw[7]));
      sts16(srow + po0, make_uint4(sw[0], sw[1], sw[2], sw[3]));
      sts16(srow + po1, make_uint4(sw[4], sw[5], sw[6], sw[7]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pd_bar);              // control: products(i), then scores(i+1)
      // ---- while they run: stage and store the results of pair i-1 ----
      if (i > 0) stage_results(i - 1, vq, vk, vv);
    }
    if (n_it > 0) {
      float vq[16], vk[16], vv[16];
      mbar_wait(o_bar, (uint32_t)((n_it - 1) & 1));
      tc_fence_after();
      tmem_ld_x16(lane_base + kColDQ + cq4 * 16, vq);
      tmem_ld_x16(lane_base + kColDK + cq4 * 16, vk);
      tmem_ld_x16(lane_base + kColDV + cq4 * 16, vv);
      tmem_ld_wait();
      tc_fence_before();
      stage_results(n_it - 1, vq, vk, vv);
    }
    if (dbias != nullptr && n_it > 0 && cq4 == 0) {
      // one thread per TMEM lane: lane m of column kColBqk holds the sum of dQ column m (m < 64) or dK column m - 64,
      // lane m of column kColBkv that of dK column m (m < 64) or dV column m - 64 — over every window this CTA processed
      mbar_wait(c_bar, 0);
      tc_fence_after();
      float bqk[16], bkv[16];
      tmem_ld_x16(lane_base + kColBqk, bqk);
      tmem_ld_x16(lane_base + kColBkv, bkv);
      tmem_ld_wait();
      tc_fence_before();
      const int c = h * HD + ri;
      atomicAdd(dbias + (own == 0 ? 0 : kv_off) + c, bqk[0]);           // dQ | dK
      if (own == 1) atomicAdd(dbias + 2 * kv_off + c, bkv[0]);          // dV
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------ forward -----------------------------------------------------------
// Same construction as the backward: two windows of one head = one 128-row problem, S = Q K^T (128 x 128 x 64) and
// O = P V (128 x 64 x 128) on tcgen05 with P written block-diagonal (shared zero half), softmax by four threads per row,
// 16 compute warps + 4 issuing warps (loader | S | O | storer), software-pipelined over the (pair, head) items of a CTA.
constexpr int kFwdControlWarps = 4;
constexpr int kFwdThreads = kComputeThreads + 32 * kFwdControlWarps;
// 2 stages x {Q, K, V} + P + O staging + exchange (max, sum) + barriers
constexpr int kFwdSmem = 1024 + 2 * 3 * kTile + kPD + kTile + 2 * 128 * 4 * 4 + 256;
constexpr uint32_t kColFS = 0, kColFO = 128;

__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, int T, int H, int64_t n_win,
                   float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // stage st: Q | K | V at smem + st * 3 * kTile
  uint8_t* Ps = smem + 6 * kTile;                     // [P_a | shared zero block | P_b]
  uint8_t* Go = Ps + kPD;                             // O staging tile
  float* xch = reinterpret_cast<float*>(Go + kTile);  // [2][128 rows][4 quarters]: row max, row sum
  uint64_t* qk_bar = reinterpret_cast<uint64_t*>(xch + 2 * 128 * 4);   // [2] Q, K of a stage landed
  uint64_t* v_bar = qk_bar + 2;                       // [2] V of a stage landed
  uint64_t* s_bar = v_bar + 2;                        // S complete
  uint64_t* sc_bar = s_bar + 1;                       // S copied to registers
  uint64_t* pd_bar = sc_bar + 1;                      // P in shared memory, previous O in registers
  uint64_t* o_bar = pd_bar + 1;                       // O complete
  uint64_t* stg_bar = o_bar + 1;                      // O staged
  uint64_t* sf_bar = stg_bar + 1;                     // staging tile free again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sf_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_pairs = (n_win + 1) >> 1;
  const int64_t n_items = n_pairs * H;                // item = pair * H + head
  const int n_it = blockIdx.x < n_items ? (int)((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  for (int i = tid; i < (6 * kTile + kPD + kTile) / 16; i += kFwdThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmO);
    for (int i = 0; i < 2; ++i) { mbar_init(&qk_bar[i], 1); mbar_init(&v_bar[i], 1); }
    mbar_init(s_bar, 1);
    mbar_init(sc_bar, kComputeThreads / 32);
    mbar_init(pd_bar, kComputeThreads / 32);
    mbar_init(o_bar, 1);
    mbar_init(stg_bar, kComputeThreads / 32);
    mbar_init(sf_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_pair = [&](int i) { return ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) / H; };
  auto item_head = [&](int i) { return (int)(((int64_t)blockIdx.x + (int64_t)i * gridDim.x) % H); };

  if (warp >= kComputeThreads / 32) {
    const int role = warp - kComputeThreads / 32;     // 0 loader, 1 S, 2 O, 3 storer
    if (lane == 0 && n_it > 0) {
      auto window_row = [&](int i, int it) {
        int64_t w = 2 * item_pair(i) + it;
        if (w >= n_win) w = n_win - 1;
        return (int32_t)(w * T);
      };
      constexpr uint64_t kTileStep = kTile >> 4, kStageStep = (3 * kTile) >> 4, kChunkStep = kChunk >> 4;
      if (role == 0) {
        const uint32_t tx_qk = 4u * (uint32_t)T * 128u, tx_v = 2u * (uint32_t)T * 128u;
        auto load_qk = [&](int i) {
          uint8_t* base = smem + (i & 1) * 3 * kTile;
          const int h = item_head(i);
          mbar_arrive_expect_tx(&qk_bar[i & 1], tx_qk);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int32_t row0 = window_row(i, it);
            tma_load_2d(base + it * 64 * 128, &tmQ, &qk_bar[i & 1], h * HD, row0);
            tma_load_2d(base + kTile + it * 64 * 128, &tmK, &qk_bar[i & 1], h * HD, row0);
          }
        };
        auto load_v = [&](int i) {
          uint8_t* base = smem + (i & 1) * 3 * kTile + 2 * kTile;
          const int h = item_head(i);
          mbar_arrive_expect_tx(&v_bar[i & 1], tx_v);
#pragma unroll
          for (int it = 0; it < 2; ++it) tma_load_2d(base + it * 64 * 128, &tmV, &v_bar[i & 1], h * HD, window_row(i, it));
        };
        auto prefetch_item = [&](int i) {
          const int h = item_head(i);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int32_t row0 = window_row(i, it);
            tma_prefetch_l2_2d(&tmQ, h * HD, row0);
            tma_prefetch_l2_2d(&tmK, h * HD, row0);
            tma_prefetch_l2_2d(&tmV, h * HD, row0);
          }
        };
        constexpr int kAhead = 4;
        load_qk(0); load_v(0);
        if (n_it > 1) { load_qk(1); load_v(1); }
        for (int i = 2; i < kAhead && i < n_it; ++i) prefetch_item(i);
        for (int i = 0; i < n_it; ++i) {
          if (i + kAhead < n_it) prefetch_item(i + kAhead);
          if (i + 2 < n_it) {
            // the stage of item i is free once O(i) is complete (its scores finished long before).  Only o_bar is waited
            // on: the loader sits on it before it can complete, whereas S may run two items ahead of a late waiter and
            // a parity wait would then miss its phase
            mbar_wait(o_bar, (uint32_t)(i & 1));
            load_qk(i + 2);
            load_v(i + 2);
          }
        }
      } else if (role == 1) {
        // ---- S(i) = Q K^T: issued as soon as S(i-1) has been copied out of TMEM ----
        const uint64_t dK_Q0 = make_smem_desc_sw128(smem_u32(smem), 0, 1024);
        const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
        for (int i = 0; i < n_it; ++i) {
          mbar_wait(&qk_bar[i & 1], (uint32_t)((i >> 1) & 1));
          if (i > 0) mbar_wait(sc_bar, (uint32_t)((i - 1) & 1));
          tc_fence_after();
          const uint64_t q = dK_Q0 + (i & 1) * kStageStep;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + kColFS, q + 2 * k, q + kTileStep + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(s_bar);
        }
      } else if (role == 2) {
        // ---- O(i) = P V: A = P K-major (K = 128 keys = two chunks 8 KB apart), B = V MN-major ----
        const uint64_t dK_P = make_smem_desc_sw128(smem_u32(Ps), 0, 1024);
        const uint64_t dM_V0 = make_smem_desc_sw128(smem_u32(smem + 2 * kTile), kTile, 1024);
        const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
        for (int i = 0; i < n_it; ++i) {
          mbar_wait(&v_bar[i & 1], (uint32_t)((i >> 1) & 1));
          mbar_wait(pd_bar, (uint32_t)(i & 1));       // P(i) written, O(i-1) in registers
          tc_fence_after();
          const uint64_t vm = dM_V0 + (i & 1) * kStageStep;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + kColFO, dK_P + (k >> 2) * kChunkStep + (k & 3) * 2, vm + 128 * k, idesc_o, k > 0 ? 1u : 0u);
          umma_commit(o_bar);
        }
      } else {
        // ---- storer ----
        for (int i = 0; i < n_it; ++i) {
          mbar_wait(stg_bar, (uint32_t)(i & 1));
          const int h = item_head(i);
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int64_t w = 2 * item_pair(i) + it;
            if (w >= n_win) break;
            tma_store_2d(&tmO, Go + it * 64 * 128, h * HD, (int32_t)(w * T));
          }
          tma_commit_group();
          tma_wait_group_read<0>();
          mbar_arrive(sf_bar);
        }
        tma_wait_group<0>();
      }
    }
  } else {
    const int q = warp & 3, cq4 = warp >> 2;
    const int r = q * 32 + lane, ri = r & 63, own = r >> 6, rsw = r & 7;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float scale_log2 = scale * 1.4426950408889634f;
    const uint32_t xrow = smem_u32(xch) + r * 16;
    const uint32_t prow = smem_u32(Ps) + own * kChunk + r * 128;
    const uint32_t po0 = (uint32_t)(((cq4 * 2) ^ rsw) << 4), po1 = (uint32_t)(((cq4 * 2 + 1) ^ rsw) << 4);
    auto stage_o = [&](int i, const float (&vo)[16]) {
      if (i > 0) mbar_wait(sf_bar, (uint32_t)((i - 1) & 1));
      if (ri < T) {
        const uint32_t ro = smem_u32(Go) + r * 128;
        sts16(ro + po0, make_uint4(pack_bf16x2(vo[0], vo[1]), pack_bf16x2(vo[2], vo[3]), pack_bf16x2(vo[4], vo[5]), pack_bf16x2(vo[6], vo[7])));
        sts16(ro + po1, make_uint4(pack_bf16x2(vo[8], vo[9]), pack_bf16x2(vo[10], vo[11]), pack_bf16x2(vo[12], vo[13]), pack_bf16x2(vo[14], vo[15])));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(stg_bar);
    };
    for (int i = 0; i < n_it; ++i) {
      mbar_wait(s_bar, (uint32_t)(i & 1));
      tc_fence_after();
      float s[16];
      tmem_ld_x16(lane_base + kColFS + own * 64 + cq4 * 16, s);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sc_bar);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (cq4 * 16 + j >= T) s[j] = -INFINITY;
        mx = fmaxf(mx, s[j]);
      }
      const float off = mx == -INFINITY ? 0.f : mx * scale_log2;
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s[j] = ex2(fmaf(s[j], scale_log2, -off));
        l += s[j];
      }
      sts_f(xrow + cq4 * 4, mx);
      sts_f(xrow + 2048 + cq4 * 4, l);
      named_bar_sync(1 + q, 128);
      float pscale;
      {
        const float4 m4 = lds_f4(xrow), l4 = lds_f4(xrow + 2048);
        const float Ml = fmaxf(fmaxf(m4.x, m4.y), fmaxf(m4.z, m4.w)) * scale_log2;
        const float f0 = ex2(fmaf(m4.x, scale_log2, -Ml)), f1 = ex2(fmaf(m4.y, scale_log2, -Ml));
        const float f2 = ex2(fmaf(m4.z, scale_log2, -Ml)), f3 = ex2(fmaf(m4.w, scale_log2, -Ml));
        l = fmaf(l4.x, f0, fmaf(l4.y, f1, fmaf(l4.z, f2, l4.w * f3)));
        pscale = (ri < T ? 1.f / l : 0.f) * (cq4 == 0 ? f0 : (cq4 == 1 ? f1 : (cq4 == 2 ? f2 : f3)));
      }
      uint32_t pw[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) pw[e] = pack_bf16x2(s[2 * e] * pscale, s[2 * e + 1] * pscale);
      float vo[16];
      if (i > 0) {
        mbar_wait(o_bar, (uint32_t)((i - 1) & 1));    // O(i-1) complete: fetch it; P tile is free
        tc_fence_after();
        tmem_ld_x16(lane_base + kColFO + cq4 * 16, vo);
        tmem_ld_wait();
      }
      tc_fence_before();
      sts16(prow + po0, make_uint4(pw[0], pw[1], pw[2], pw[3]));
      sts16(prow + po1, make_uint4(pw[4], pw[5], pw[6], pw[7]));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pd_bar);
      if (i > 0) stage_o(i - 1, vo);
      // the exchange slots are rewritten in the next iteration only after every partner has passed this iteration's
      // named barrier and (through sc_bar -> s_bar) read them
    }
    if (n_it > 0) {
      float vo[16];
      mbar_wait(o_bar, (uint32_t)((n_it - 1) & 1));
      tc_fence_after();
      tmem_ld_x16(lane_base + kColFO + cq4 * 16, vo);
      tmem_ld_wait();
      tc_fence_before();
      stage_o(n_it - 1, vo);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace attn_tc

int attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                     int64_t n_win, int T, int H, int head_dim, float scale, cudaStream_t s) {
  using namespace attn_tc;
  if (head_dim != HD || T > 64 || T < 1 || scale <= 0.f || n_win * T >= (1ll << 31)) return IBM_E_UNSUPPORTED;
  const int64_t rows = n_win * T, cols = (int64_t)H * HD;
  CUtensorMap mq, mk, mv, mo;
  int rc = make_map(&mq, q, false, cols, rows, ldq, 64, (uint32_t)T);
  if (rc) return rc;
  rc = make_map(&mk, k, false, cols, rows, ldk, 64, (uint32_t)T);
  if (rc) return rc;
  rc = make_map(&mv, v, false, cols, rows, ldv, 64, (uint32_t)T);
  if (rc) return rc;
  rc = make_map(&mo, o, false, cols, rows, ldo, 64, (uint32_t)T);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    IBM_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem));
    attr_set = true;
  }
  const int64_t n_items = (n_win + 1) / 2 * H;
  const int64_t grid = n_items < sm_count() ? n_items : sm_count();
  attn_fwd_tc_kernel<<<(unsigned)grid, kFwdThreads, kFwdSmem, s>>>(mq, mk, mv, mo, T, H, n_win, scale);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

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
  attn_bwd_tc_kernel<<<(unsigned)grid, kThreads, kSmem, s>>>(mq, md, mg, T, H, n_win, kv_off, scale, dbias,
                                                             next_walk_reverse(n_win * T * (int64_t)H * HD * 2 * 7));
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

}  // namespace ibm
