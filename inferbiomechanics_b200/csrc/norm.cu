// LayerNorm forward/backward over bf16 rows (fp32 statistics), warp-per-row with shuffle
// reductions.  Replaces nn.LayerNorm at /root/reference/src/models/TransformerBaseline.py:21-22,31,36
// (the residual add is fused into the producing GEMM's epilogue, see gemm_sm100.cu aux_mode 1).
// HBM-bound: forward reads s and writes y (2 units); backward reads dy, s and writes ds (3 units).
#include <stdlib.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace ibm {

using namespace ptx;

// Rows stream through shared memory: every warp owns a private ring of kStages row groups filled by bulk async
// copies (cp.async.bulk → mbarrier complete_tx, issued by lane 0) that run kStages-1 groups ahead of the warp's
// arithmetic.  The bytes in flight per SM no longer depend on registers per thread (the register-resident
// predecessors of these kernels sat at 45-50 % of the HBM roofline with 2 blocks per SM; profiles/r01a, r01b).
constexpr int kThreads = 256;          // 8 warps, each an independent pipeline
constexpr int kWarps = kThreads / 32;
constexpr int kMaxChunks = 4;          // row length <= 4 * 32 lanes * 8 = 1024 columns
constexpr int kStages = 4;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint4 lds16(const void* p) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
  return r;
}
__device__ __forceinline__ void unpack8(uint4 u, float (&v)[8]) {
  float2 a;
  a = unpack_bf16x2(u.x); v[0] = a.x; v[1] = a.y;
  a = unpack_bf16x2(u.y); v[2] = a.x; v[3] = a.y;
  a = unpack_bf16x2(u.z); v[4] = a.x; v[5] = a.y;
  a = unpack_bf16x2(u.w); v[6] = a.x; v[7] = a.y;
}

// Packed fp32 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per instruction and issue slot).  The
// LayerNorm kernels are issue-bound at the clock the power-capped training step runs at (ncu, profiles/r01d: 59 % / 71 %
// issue-active at 1.9 GHz for forward / backward, half of the instructions elementwise fp32 math on 16 values per lane):
// pairing the elementwise work takes the forward from ~222 to ~125 and the backward from ~402 to ~200 warp instructions
// per row.  Each lane of a pair is rounded exactly like the scalar instruction it replaces.
__device__ __forceinline__ void unpack8p(uint4 u, float2 (&v)[4]) {
  v[0] = unpack_bf16x2(u.x); v[1] = unpack_bf16x2(u.y); v[2] = unpack_bf16x2(u.z); v[3] = unpack_bf16x2(u.w);
}
__device__ __forceinline__ uint4 pack8p(const float2 (&o)[4]) {
  return make_uint4(pack_bf16x2(o[0].x, o[0].y), pack_bf16x2(o[1].x, o[1].y), pack_bf16x2(o[2].x, o[2].y), pack_bf16x2(o[3].x, o[3].y));
}

// each lane owns chunks of 8 consecutive columns: columns (k*32 + lane)*8 … +7 for k < CH.  A row group is R
// consecutive rows (R * ld * 2 contiguous bytes: one bulk copy).
// FULL: d == ld == CH*256 (no pad columns, every chunk complete) — the column predicates fold away, which halves
// the instruction count of these issue-bound kernels (profiles/r01c: 355 → ~190 warp instructions per row forward).
template <int CH, int R, bool FULL, bool PK>
__global__ void __launch_bounds__(kThreads)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ y, long long ld,
                     const float* __restrict__ gamma, const float* __restrict__ beta, long long M, int d, float eps,
                     float* __restrict__ mean, float* __restrict__ rstd, int rev) {
  extern __shared__ __align__(128) uint8_t smem_ln[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)ld * 2u;
  const uint32_t group_bytes = R * row_bytes;
  // layout: [gamma d fp32][beta d fp32][8 warps][kStages][R rows] then the mbarriers
  float* sg = reinterpret_cast<float*>(smem_ln);
  float* sb = sg + CH * 256;
  uint8_t* ring = reinterpret_cast<uint8_t*>(sb + CH * 256) + (size_t)wid * kStages * group_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sb + CH * 256) + (size_t)kWarps * kStages * group_bytes) + wid * kStages;
  for (int c = threadIdx.x; c < CH * 256; c += kThreads) {
    sg[c] = c < d ? __ldg(gamma + c) : 0.f;
    sb[c] = c < d ? __ldg(beta + c) : 0.f;
  }
  if (lane == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const long long n_groups = (M + R - 1) / R;
  const long long gw = (long long)blockIdx.x * kWarps + wid, gstride = (long long)gridDim.x * kWarps;
  // rev: the grid walks the row groups from the last one down (ibm_set_walk_order)
  auto issue = [&](long long grp, int st) {        // lane 0
    const long long m0 = (rev ? n_groups - 1 - grp : grp) * R;
    const uint32_t bytes = (uint32_t)((M - m0 < R ? M - m0 : R)) * row_bytes;
    mbar_arrive_expect_tx(&bars[st], bytes);
    bulk_load(ring + (size_t)st * group_bytes, s + m0 * ld, bytes, &bars[st]);
  };
  if (lane == 0) {
    for (int i = 0; i < kStages; ++i)
      if (gw + i * gstride < n_groups) issue(gw + i * gstride, i);
  }
  const float inv_d = 1.f / (float)d;
  // full rows of up to 512 columns: gamma / beta of this lane's 16 columns live in registers (read from shared memory, the
  // 32-byte lane stride of the float4 reads costs two wavefronts each: 64 of the 72 LSU wavefronts per row were these)
  constexpr bool kRegGB = FULL && PK && CH <= 2;
  float2 G2[kRegGB ? CH : 1][4], B2[kRegGB ? CH : 1][4];
  if constexpr (kRegGB) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (k * 32 + lane) * 8 + 2 * i;
        G2[k][i] = make_float2(__ldg(gamma + c), __ldg(gamma + c + 1));
        B2[k][i] = make_float2(__ldg(beta + c), __ldg(beta + c + 1));
      }
  }
  int it = 0;
  for (long long grp = gw; grp < n_groups; grp += gstride, ++it) {
    const int st = it % kStages;
    mbar_wait(&bars[st], (uint32_t)((it / kStages) & 1));
    const uint8_t* tile = ring + (size_t)st * group_bytes;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long m = (rev ? n_groups - 1 - grp : grp) * R + r;
      if (m >= M) break;
      if constexpr (FULL && PK) {
        float2 v[CH][4];
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          unpack8p(lds16(tile + r * row_bytes + (k * 32 + lane) * 16), v[k]);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc = __fadd2_rn(acc, v[k][i]);
        }
        const float mu = warp_sum(acc.x + acc.y) * inv_d;
        const float2 nmu = make_float2(-mu, -mu);
        acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CH; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[k][i] = __fadd2_rn(v[k][i], nmu);
            acc = __ffma2_rn(v[k][i], v[k][i], acc);
          }
        const float rs = rsqrtf(warp_sum(acc.x + acc.y) * inv_d + eps);
        const float2 rs2 = make_float2(rs, rs);
        if (lane == 0) { if (mean) mean[m] = mu; if (rstd) rstd[m] = rs; }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const int c = (k * 32 + lane) * 8;
          float2 o[4];
          if constexpr (kRegGB) {
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = __ffma2_rn(__fmul2_rn(v[k][i], rs2), G2[k][i], B2[k][i]);
          } else {
            const float4 g0 = *reinterpret_cast<const float4*>(sg + c), g1 = *reinterpret_cast<const float4*>(sg + c + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(sb + c), b1 = *reinterpret_cast<const float4*>(sb + c + 4);
            o[0] = __ffma2_rn(__fmul2_rn(v[k][0], rs2), make_float2(g0.x, g0.y), make_float2(b0.x, b0.y));
            o[1] = __ffma2_rn(__fmul2_rn(v[k][1], rs2), make_float2(g0.z, g0.w), make_float2(b0.z, b0.w));
            o[2] = __ffma2_rn(__fmul2_rn(v[k][2], rs2), make_float2(g1.x, g1.y), make_float2(b1.x, b1.y));
            o[3] = __ffma2_rn(__fmul2_rn(v[k][3], rs2), make_float2(g1.z, g1.w), make_float2(b1.z, b1.w));
          }
          st_stream16(y + m * ld + c, pack8p(o));
        }
        continue;
      }
      float v[CH][8];
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (FULL || c < ld) {
          unpack8(lds16(tile + r * row_bytes + c * 2), v[k]);
#pragma unroll
          for (int j = 0; j < 8; ++j) { if (!FULL && c + j >= d) v[k][j] = 0.f; sum += v[k][j]; }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
        }
      }
      const float mu = warp_sum(sum) * inv_d;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (FULL || c + j < d) { float t = v[k][j] - mu; sq = fmaf(t, t, sq); }
      }
      const float rs = rsqrtf(warp_sum(sq) * inv_d + eps);
      if (lane == 0) { if (mean) mean[m] = mu; if (rstd) rstd[m] = rs; }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (FULL || c < ld) {
          const float4 g0 = *reinterpret_cast<const float4*>(sg + c), g1 = *reinterpret_cast<const float4*>(sg + c + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sb + c), b1 = *reinterpret_cast<const float4*>(sb + c + 4);
          const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (FULL || c + j < d) ? fmaf((v[k][j] - mu) * rs, gm[j], bt[j]) : 0.f;
          st_stream16(y + m * ld + c, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                 pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
        }
      }
    }
    __syncwarp();                                   // every lane has read this stage
    if (lane == 0 && grp + kStages * gstride < n_groups) issue(grp + kStages * gstride, st);
  }
}

// Narrow rows (ld <= 128, e.g. the TransformerBaseline's d = 108 padded to 112 columns; TransformerBaseline.py:79,87):
// a warp per row would leave 18 of 32 lanes idle and pay two 5-step shuffle reductions per 224-byte row.  Here 8 lanes
// share a row (lane sub owns the 16-byte chunks sub and 8 + sub, so a quarter-warp reads 128 contiguous bytes of shared
// memory: no bank conflicts), a warp works on 4 rows per instruction, reductions are 3 shuffle steps, and gamma/beta of a
// lane's 16 columns live in registers for the whole kernel.  Row groups of RN rows stream through the same per-warp
// bulk-copy ring as the wide kernel.
template <int RN>
__global__ void __launch_bounds__(kThreads)
layernorm_fwd_narrow_kernel(const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ y, long long ld,
                            const float* __restrict__ gamma, const float* __restrict__ beta, long long M, int d, float eps,
                            float* __restrict__ mean, float* __restrict__ rstd) {
  extern __shared__ __align__(128) uint8_t smem_ln[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int sub = lane & 7, rr = lane >> 3;
  const uint32_t row_bytes = (uint32_t)ld * 2u;
  const uint32_t group_bytes = RN * row_bytes;
  uint8_t* ring = smem_ln + (size_t)wid * kStages * group_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ln + (size_t)kWarps * kStages * group_bytes) + wid * kStages;
  if (lane == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  const int c0 = sub * 8, c1 = 64 + sub * 8;
  const bool has0 = c0 < ld, has1 = c1 < ld;         // ld < 64: the upper lanes of a row group own no chunk at all
  float gm[16], bt[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    gm[j] = c0 + j < d ? __ldg(gamma + c0 + j) : 0.f;
    bt[j] = c0 + j < d ? __ldg(beta + c0 + j) : 0.f;
    gm[8 + j] = c1 + j < d ? __ldg(gamma + c1 + j) : 0.f;
    bt[8 + j] = c1 + j < d ? __ldg(beta + c1 + j) : 0.f;
  }
  const long long n_groups = (M + RN - 1) / RN;
  const long long gw = (long long)blockIdx.x * kWarps + wid, gstride = (long long)gridDim.x * kWarps;
  auto issue = [&](long long grp, int st) {        // lane 0
    const long long m0 = grp * RN;
    const uint32_t bytes = (uint32_t)((M - m0 < RN ? M - m0 : RN)) * row_bytes;
    mbar_arrive_expect_tx(&bars[st], bytes);
    bulk_load(ring + (size_t)st * group_bytes, s + m0 * ld, bytes, &bars[st]);
  };
  if (lane == 0) {
    for (int i = 0; i < kStages; ++i)
      if (gw + i * gstride < n_groups) issue(gw + i * gstride, i);
  }
  const float inv_d = 1.f / (float)d;
  int it = 0;
  for (long long grp = gw; grp < n_groups; grp += gstride, ++it) {
    const int st = it % kStages;
    mbar_wait(&bars[st], (uint32_t)((it / kStages) & 1));
    const uint8_t* tile = ring + (size_t)st * group_bytes;
#pragma unroll
    for (int r4 = 0; r4 < RN; r4 += 4) {
      const int r = r4 + rr;
      const long long m = grp * RN + r;
      const bool live = m < M;                       // rows past M hold stale shared memory: computed, never stored
      float va[8], vb[8], v[16];
      if (has0) unpack8(lds16(tile + r * row_bytes + c0 * 2), va);
      if (has1) unpack8(lds16(tile + r * row_bytes + c1 * 2), vb);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = has0 && c0 + j < d ? va[j] : 0.f;
        v[8 + j] = has1 && c1 + j < d ? vb[j] : 0.f;
        sum += v[j] + v[8 + j];
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mu = sum * inv_d;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j < d) { const float t = v[j] - mu; sq = fmaf(t, t, sq); }
        if (has1 && c1 + j < d) { const float t = v[8 + j] - mu; sq = fmaf(t, t, sq); }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rs = rsqrtf(sq * inv_d + eps);
      if (live) {
        if (sub == 0) { if (mean) mean[m] = mu; if (rstd) rstd[m] = rs; }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaf((v[j] - mu) * rs, gm[j], bt[j]);     // gm = bt = 0 beyond d: pads get 0
        if (has0)
          st_stream16(y + m * ld + c0, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                  pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
        if (has1)
          st_stream16(y + m * ld + c1, make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]),
                                                  pack_bf16x2(o[12], o[13]), pack_bf16x2(o[14], o[15])));
      }
    }
    __syncwarp();                                   // every lane has read this stage
    if (lane == 0 && grp + kStages * gstride < n_groups) issue(grp + kStages * gstride, st);
  }
}

// Backward.  ds = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma,  xhat = (s-mu)*rstd.
// Column reductions (dgamma, dbeta, colsum(ds)) are accumulated in registers over the rows a warp
// visits, combined across the block's 8 warps in shared memory, then one fp32 atomic per column
// per block.  A ring stage holds one row of dy and one row of s.
template <int CH, bool FULL, bool PK>
__global__ void __launch_bounds__(kThreads, CH <= 2 ? 2 : 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ s, long long ld,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     long long M, int d, __nv_bfloat16* __restrict__ ds, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dcolsum, int rev) {
  extern __shared__ __align__(128) uint8_t smem_ln[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)ld * 2u;
  // layout: [gamma CH*256 fp32][8 warps][kStages][dy row | s row] then mbarriers; the ring is re-used for the final
  // block combine (needs 3 * 8 * CH*256 fp32 = 24 KB * CH, the ring is 8 * kStages * 2 * row_bytes >= that for kStages >= 3
  // only when rows are full: the launcher sizes the allocation as the max of the two)
  float* sg = reinterpret_cast<float*>(smem_ln);
  uint8_t* ring_all = reinterpret_cast<uint8_t*>(sg + CH * 256);
  uint8_t* ring = ring_all + (size_t)wid * kStages * 2 * row_bytes;
  const size_t ring_bytes = (size_t)kWarps * kStages * 2 * row_bytes;
  const size_t comb_bytes = (size_t)3 * kWarps * CH * 256 * sizeof(float);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_all + (ring_bytes > comb_bytes ? ring_bytes : comb_bytes)) + wid * kStages;
  for (int c = threadIdx.x; c < CH * 256; c += kThreads) sg[c] = c < d ? __ldg(gamma + c) : 0.f;
  if (lane == 0) {
    for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();

  const long long gw = (long long)blockIdx.x * kWarps + wid, gstride = (long long)gridDim.x * kWarps;
  auto phys = [&](long long m) { return rev ? M - 1 - m : m; };      // rev: rows are visited from the last one down
  auto issue = [&](long long mv, int st) {          // lane 0
    const long long m = phys(mv);
    mbar_arrive_expect_tx(&bars[st], 2 * row_bytes);
    uint8_t* dst = ring + (size_t)st * 2 * row_bytes;
    bulk_load(dst, dy + m * ld, row_bytes, &bars[st]);
    bulk_load(dst + row_bytes, s + m * ld, row_bytes, &bars[st]);
  };
  if (lane == 0) {
    for (int i = 0; i < kStages; ++i)
      if (gw + i * gstride < M) issue(gw + i * gstride, i);
  }
  const float inv_d = 1.f / (float)d;
  // column accumulators as fp32 pairs (see the note on packed arithmetic above); element j of chunk k is half (j & 1) of pair j / 2
  float2 ag2[CH][4], ab2[CH][4], ac2[CH][4];
#define LN_ACC(a, k, j) (((j) & 1) ? a[k][(j) >> 1].y : a[k][(j) >> 1].x)
#pragma unroll
  for (int k = 0; k < CH; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) ag2[k][i] = ab2[k][i] = ac2[k][i] = make_float2(0.f, 0.f);
  constexpr bool kRegG = FULL && PK && CH <= 2;      // gamma of this lane's 16 columns in registers
  float2 GM[kRegG ? CH : 1][4];
  if constexpr (kRegG) {
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (k * 32 + lane) * 8 + 2 * i;
        GM[k][i] = make_float2(__ldg(gamma + c), __ldg(gamma + c + 1));
      }
  }
  int it = 0;
  float mu_n = gw < M ? __ldg(mean + phys(gw)) : 0.f, rs_n = gw < M ? __ldg(rstd + phys(gw)) : 0.f;
  for (long long m = gw; m < M; m += gstride, ++it) {
    const int st = it % kStages;
    const float mu = mu_n, rs = rs_n;
    if (m + gstride < M) { mu_n = __ldg(mean + phys(m + gstride)); rs_n = __ldg(rstd + phys(m + gstride)); }   // next row's statistics
    mbar_wait(&bars[st], (uint32_t)((it / kStages) & 1));
    const uint8_t* tile = ring + (size_t)st * 2 * row_bytes;
    if constexpr (FULL && PK) {
      float2 xh[CH][4], g[CH][4];
      float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
      const float2 nmu = make_float2(-mu, -mu), rs2 = make_float2(rs, rs);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
        unpack8p(lds16(tile + c * 2), g[k]);
        unpack8p(lds16(tile + row_bytes + c * 2), xh[k]);
        float2 gm[4];
        if constexpr (kRegG) {
#pragma unroll
          for (int i = 0; i < 4; ++i) gm[i] = GM[k][i];
        } else {
          const float4 g0 = *reinterpret_cast<const float4*>(sg + c), g1 = *reinterpret_cast<const float4*>(sg + c + 4);
          gm[0] = make_float2(g0.x, g0.y); gm[1] = make_float2(g0.z, g0.w); gm[2] = make_float2(g1.x, g1.y); gm[3] = make_float2(g1.z, g1.w);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          xh[k][i] = __fmul2_rn(__fadd2_rn(xh[k][i], nmu), rs2);
          ag2[k][i] = __ffma2_rn(g[k][i], xh[k][i], ag2[k][i]);       // dgamma
          ab2[k][i] = __fadd2_rn(ab2[k][i], g[k][i]);                  // dbeta
          g[k][i] = __fmul2_rn(g[k][i], gm[i]);
          a1 = __fadd2_rn(a1, g[k][i]);
          a2 = __ffma2_rn(g[k][i], xh[k][i], a2);
        }
      }
      __syncwarp();                                   // every lane has read this stage
      if (lane == 0 && m + kStages * gstride < M) issue(m + kStages * gstride, st);
      const float s1 = warp_sum(a1.x + a1.y) * inv_d, s2 = warp_sum(a2.x + a2.y) * inv_d;
      const float2 ns1 = make_float2(-s1, -s1), ns2 = make_float2(-s2, -s2);
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        float2 o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[i] = __fmul2_rn(__ffma2_rn(xh[k][i], ns2, __fadd2_rn(g[k][i], ns1)), rs2);
          ac2[k][i] = __fadd2_rn(ac2[k][i], o[i]);
        }
        st_stream16(ds + phys(m) * ld + (k * 32 + lane) * 8, pack8p(o));
      }
      continue;
    }
    float xh[CH][8], g[CH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (FULL || c < ld) {
        unpack8(lds16(tile + c * 2), g[k]);
        unpack8(lds16(tile + row_bytes + c * 2), xh[k]);
        const float4 g0 = *reinterpret_cast<const float4*>(sg + c), g1 = *reinterpret_cast<const float4*>(sg + c + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (FULL || c + j < d) {
            xh[k][j] = (xh[k][j] - mu) * rs;
            LN_ACC(ag2, k, j) = fmaf(g[k][j], xh[k][j], LN_ACC(ag2, k, j));     // dgamma
            LN_ACC(ab2, k, j) += g[k][j];                                         // dbeta
            g[k][j] *= gm[j];
            s1 += g[k][j];
            s2 = fmaf(g[k][j], xh[k][j], s2);
          } else { xh[k][j] = 0.f; g[k][j] = 0.f; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { xh[k][j] = 0.f; g[k][j] = 0.f; }
      }
    }
    __syncwarp();                                   // every lane has read this stage
    if (lane == 0 && m + kStages * gstride < M) issue(m + kStages * gstride, st);
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (FULL || c < ld) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = (FULL || c + j < d) ? rs * (g[k][j] - s1 - xh[k][j] * s2) : 0.f;
          LN_ACC(ac2, k, j) += o[j];
        }
        st_stream16(ds + phys(m) * ld + c, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                      pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
      }
    }
  }
  // block combine: sm[q][wid][col] over the (now idle) ring
  __syncthreads();
  float* sm = reinterpret_cast<float*>(ring_all);
  constexpr int W = CH * 256;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = (k * 32 + lane) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[(0 * 8 + wid) * W + c + j] = LN_ACC(ag2, k, j);
      sm[(1 * 8 + wid) * W + c + j] = LN_ACC(ab2, k, j);
      sm[(2 * 8 + wid) * W + c + j] = LN_ACC(ac2, k, j);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += kThreads) {
    float a = 0.f, b = 0.f, e = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += sm[(0 * 8 + w) * W + c]; b += sm[(1 * 8 + w) * W + c]; e += sm[(2 * 8 + w) * W + c]; }
    if (dgamma) atomicAdd(dgamma + c, a);
    if (dbeta) atomicAdd(dbeta + c, b);
    if (dcolsum) atomicAdd(dcolsum + c, e);
  }
#undef LN_ACC
}

// persistent grid: blocks per SM from the occupancy calculator, capped by the work
template <typename K>
static int ln_grid(K kern, size_t smem, long long units, int* grid) {
  int per_sm = 0;
  IBM_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  long long cap = (long long)sm_count() * per_sm;
  long long need = ceil_div(units, kWarps);
  *grid = (int)(need < 1 ? 1 : (need < cap ? need : cap));
  return IBM_OK;
}

// IBM_LN_PACKED=0: the scalar arithmetic of round 1 (A/B measurements)
static bool ln_packed() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("IBM_LN_PACKED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int CH, int R, bool FULL, bool PK = true>
static int launch_ln_fwd(const __nv_bfloat16* sp, __nv_bfloat16* yp, int64_t ld, const float* gamma, const float* beta, int64_t M,
                         int32_t d, float eps, float* mean, float* rstd, cudaStream_t st) {
  if constexpr (FULL && PK) {
    if (!ln_packed()) return launch_ln_fwd<CH, R, FULL, false>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd, st);
  }
  auto kern = layernorm_fwd_kernel<CH, R, FULL, FULL && PK>;
  const size_t smem = (size_t)2 * CH * 256 * sizeof(float) + (size_t)kWarps * kStages * R * ld * 2 + kWarps * kStages * 8;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    IBM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int grid = 0;
  int rc = ln_grid(kern, smem, ceil_div(M, R), &grid);
  if (rc) return rc;
  kern<<<grid, kThreads, smem, st>>>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd, next_walk_reverse(2 * M * ld * 2));
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

template <int RN>
static int launch_ln_fwd_narrow(const __nv_bfloat16* sp, __nv_bfloat16* yp, int64_t ld, const float* gamma, const float* beta,
                                int64_t M, int32_t d, float eps, float* mean, float* rstd, cudaStream_t st) {
  auto kern = layernorm_fwd_narrow_kernel<RN>;
  const size_t smem = (size_t)kWarps * kStages * RN * ld * 2 + kWarps * kStages * 8;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    IBM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int grid = 0;
  int rc = ln_grid(kern, smem, ceil_div(M, RN), &grid);
  if (rc) return rc;
  kern<<<grid, kThreads, smem, st>>>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

template <int CH, bool FULL, bool PK = true>
static int launch_ln_bwd(const __nv_bfloat16* dyp, const __nv_bfloat16* sp, int64_t ld, const float* gamma, const float* mean,
                         const float* rstd, int64_t M, int32_t d, __nv_bfloat16* dsp, float* dgamma, float* dbeta, float* dcolsum,
                         cudaStream_t st) {
  if constexpr (FULL && PK) {
    if (!ln_packed()) return launch_ln_bwd<CH, FULL, false>(dyp, sp, ld, gamma, mean, rstd, M, d, dsp, dgamma, dbeta, dcolsum, st);
  }
  auto kern = layernorm_bwd_kernel<CH, FULL, FULL && PK>;
  const size_t ring = (size_t)kWarps * kStages * 2 * ld * 2, comb = (size_t)3 * kWarps * CH * 256 * sizeof(float);
  const size_t smem = (size_t)CH * 256 * sizeof(float) + (ring > comb ? ring : comb) + kWarps * kStages * 8;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    IBM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int grid = 0;
  int rc = ln_grid(kern, smem, M, &grid);
  if (rc) return rc;
  kern<<<grid, kThreads, smem, st>>>(dyp, sp, ld, gamma, mean, rstd, M, d, dsp, dgamma, dbeta, dcolsum,
                                     next_walk_reverse(3 * M * ld * 2));
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

}  // namespace ibm

extern "C" int ibm_layernorm_fwd(const void* s, void* y, int64_t ld, const float* gamma, const float* beta, int64_t M,
                                 int32_t d, float eps, float* mean, float* rstd, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(s && y && gamma && beta && M > 0 && d > 0, "layernorm_fwd: bad argument");
  IBM_CHECK_ARG(ld % 8 == 0 && ld >= d && ld <= kMaxChunks * 256 && aligned16(s) && aligned16(y),
                "layernorm_fwd: ld must be a multiple of 8, >= d and <= 1024; pointers 16-byte aligned");
  const int ch = (int)ceil_div(ld, 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* sp = static_cast<const __nv_bfloat16*>(s);
  auto* yp = static_cast<__nv_bfloat16*>(y);
  const bool full = d == ld && ld == (int64_t)ch * 256;
  if (ld <= 128) return launch_ln_fwd_narrow<8>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd, st);
#define IBM_LN_FWD(CHV, RV)                                                                              \
  return full ? launch_ln_fwd<CHV, RV, true>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd, st)         \
              : launch_ln_fwd<CHV, RV, false>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd, st)
  switch (ch) {
    case 1: IBM_LN_FWD(1, 4);
    case 2: IBM_LN_FWD(2, 2);
    case 3: IBM_LN_FWD(3, 1);
    default: IBM_LN_FWD(4, 1);
  }
#undef IBM_LN_FWD
}

extern "C" int ibm_layernorm_bwd(const void* dy, const void* s, int64_t ld, const float* gamma, const float* mean,
                                 const float* rstd, int64_t M, int32_t d, void* ds, float* dgamma, float* dbeta,
                                 float* dcolsum, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(dy && s && gamma && mean && rstd && ds && M > 0 && d > 0, "layernorm_bwd: bad argument");
  IBM_CHECK_ARG(ld % 8 == 0 && ld >= d && ld <= kMaxChunks * 256 && aligned16(s) && aligned16(dy) && aligned16(ds),
                "layernorm_bwd: ld must be a multiple of 8, >= d and <= 1024; pointers 16-byte aligned");
  const int ch = (int)ceil_div(ld, 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* dyp = static_cast<const __nv_bfloat16*>(dy);
  auto* sp = static_cast<const __nv_bfloat16*>(s);
  auto* dsp = static_cast<__nv_bfloat16*>(ds);
  const bool full = d == ld && ld == (int64_t)ch * 256;
#define IBM_LN_BWD(CHV)                                                                                               \
  return full ? launch_ln_bwd<CHV, true>(dyp, sp, ld, gamma, mean, rstd, M, d, dsp, dgamma, dbeta, dcolsum, st)       \
              : launch_ln_bwd<CHV, false>(dyp, sp, ld, gamma, mean, rstd, M, d, dsp, dgamma, dbeta, dcolsum, st)
  switch (ch) {
    case 1: IBM_LN_BWD(1);
    case 2: IBM_LN_BWD(2);
    case 3: IBM_LN_BWD(3);
    default: IBM_LN_BWD(4);
  }
#undef IBM_LN_BWD
}
