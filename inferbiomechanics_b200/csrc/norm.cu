// LayerNorm forward/backward over bf16 rows (fp32 statistics), warp-per-row with shuffle
// reductions.  Replaces nn.LayerNorm at /root/reference/src/models/TransformerBaseline.py:21-22,31,36
// (the residual add is fused into the producing GEMM's epilogue, see gemm_sm100.cu aux_mode 1).
// HBM-bound: forward reads s and writes y (2 units); backward reads dy, s and writes ds (3 units).
#include "common.cuh"

namespace ibm {

constexpr int kThreads = 256;          // 8 rows per block-iteration
constexpr int kMaxChunks = 4;          // row length <= 4 * 32 lanes * 8 = 1024 columns

// each lane owns chunks of 8 consecutive columns: columns (k*32 + lane)*8 … +7.  A warp handles RPW
// rows per iteration with all their loads issued before the first reduction (memory-level parallelism:
// one row per warp left the kernel latency-bound at ~30 % of HBM bandwidth, see profiles/r01a_*).
template <int CH, int RPW>
__global__ void __launch_bounds__(kThreads)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ s, __nv_bfloat16* __restrict__ y, long long ld,
                     const float* __restrict__ gamma, const float* __restrict__ beta, long long M, int d, float eps,
                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.f / (float)d;
  // gamma/beta of this lane's columns live in registers for the whole kernel (re-loading them per row made the
  // kernel LSU-bound: 32 scalar loads per row per lane, see profiles/r01a)
  float gm[CH][8], bt[CH][8];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = (k * 32 + lane) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gm[k][j] = (c + j < d) ? __ldg(gamma + c + j) : 0.f;
      bt[k][j] = (c + j < d) ? __ldg(beta + c + j) : 0.f;
    }
  }
  for (long long m0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; m0 < M; m0 += warps * RPW) {
    float v[RPW][CH][8];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const long long m = m0 + r;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
        if (m < M && c < ld) {
          uint4 u = ld_stream16(s + m * ld + c);
          float2 a;
          a = unpack_bf16x2(u.x); v[r][k][0] = a.x; v[r][k][1] = a.y;
          a = unpack_bf16x2(u.y); v[r][k][2] = a.x; v[r][k][3] = a.y;
          a = unpack_bf16x2(u.z); v[r][k][4] = a.x; v[r][k][5] = a.y;
          a = unpack_bf16x2(u.w); v[r][k][6] = a.x; v[r][k][7] = a.y;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[r][k][j] = 0.f;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const long long m = m0 + r;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) { if (c + j >= d) v[r][k][j] = 0.f; sum += v[r][k][j]; }
      }
      const float mu = warp_sum(sum) * inv_d;
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = (k * 32 + lane) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (c + j < d) { float t = v[r][k][j] - mu; sq = fmaf(t, t, sq); }
      }
      const float rs = rsqrtf(warp_sum(sq) * inv_d + eps);
      if (m < M) {
        if (lane == 0) { if (mean) mean[m] = mu; if (rstd) rstd[m] = rs; }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const int c = (k * 32 + lane) * 8;
          if (c < ld) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = (c + j < d) ? fmaf((v[r][k][j] - mu) * rs, gm[k][j], bt[k][j]) : 0.f;
            st_stream16(y + m * ld + c, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                   pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
          }
        }
      }
    }
  }
}

// Backward.  ds = rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma,  xhat = (s-mu)*rstd.
// Column reductions (dgamma, dbeta, colsum(ds)) are accumulated in registers over the rows a warp
// visits, combined across the block's 8 warps in shared memory, then one fp32 atomic per column
// per block.
template <int CH>
__global__ void __launch_bounds__(kThreads)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ s, long long ld,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     long long M, int d, __nv_bfloat16* __restrict__ ds, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dcolsum) {
  extern __shared__ float sm[];          // [3][8 warps][CH*256]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float inv_d = 1.f / (float)d;
  float ag[CH][8], ab[CH][8], ac[CH][8];
  float gm[CH][8];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = (k * 32 + lane) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { ag[k][j] = ab[k][j] = ac[k][j] = 0.f; gm[k][j] = (c + j < d) ? __ldg(gamma + c + j) : 0.f; }
  }
  for (long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    const float mu = __ldg(mean + m), rs = __ldg(rstd + m);
    float xh[CH][8], g[CH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < ld) {
        uint4 us = ld_stream16(s + m * ld + c), ud = ld_stream16(dy + m * ld + c);
        float2 a, b;
        a = unpack_bf16x2(us.x); b = unpack_bf16x2(ud.x); xh[k][0] = a.x; xh[k][1] = a.y; g[k][0] = b.x; g[k][1] = b.y;
        a = unpack_bf16x2(us.y); b = unpack_bf16x2(ud.y); xh[k][2] = a.x; xh[k][3] = a.y; g[k][2] = b.x; g[k][3] = b.y;
        a = unpack_bf16x2(us.z); b = unpack_bf16x2(ud.z); xh[k][4] = a.x; xh[k][5] = a.y; g[k][4] = b.x; g[k][5] = b.y;
        a = unpack_bf16x2(us.w); b = unpack_bf16x2(ud.w); xh[k][6] = a.x; xh[k][7] = a.y; g[k][6] = b.x; g[k][7] = b.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c + j < d) {
            xh[k][j] = (xh[k][j] - mu) * rs;
            ag[k][j] = fmaf(g[k][j], xh[k][j], ag[k][j]);     // dgamma
            ab[k][j] += g[k][j];                               // dbeta
            g[k][j] *= gm[k][j];
            s1 += g[k][j];
            s2 = fmaf(g[k][j], xh[k][j], s2);
          } else { xh[k][j] = 0.f; g[k][j] = 0.f; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { xh[k][j] = 0.f; g[k][j] = 0.f; }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < ld) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = (c + j < d) ? rs * (g[k][j] - s1 - xh[k][j] * s2) : 0.f;
          ac[k][j] += o[j];
        }
        st_stream16(ds + m * ld + c, make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7])));
      }
    }
  }
  // block combine: sm[q][wid][col]
  constexpr int W = CH * 256;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int c = (k * 32 + lane) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sm[(0 * 8 + wid) * W + c + j] = ag[k][j];
      sm[(1 * 8 + wid) * W + c + j] = ab[k][j];
      sm[(2 * 8 + wid) * W + c + j] = ac[k][j];
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
}

static int rows_grid(long long M) {
  long long need = ceil_div(M, kThreads / 32);
  long long cap = (long long)sm_count() * 8;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
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
  const int grid = rows_grid(M);
  switch (ch) {
    case 1: layernorm_fwd_kernel<1, 4><<<grid, kThreads, 0, st>>>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd); break;
    case 2: layernorm_fwd_kernel<2, 2><<<grid, kThreads, 0, st>>>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd); break;
    case 3: layernorm_fwd_kernel<3, 1><<<grid, kThreads, 0, st>>>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd); break;
    default: layernorm_fwd_kernel<4, 1><<<grid, kThreads, 0, st>>>(sp, yp, ld, gamma, beta, M, d, eps, mean, rstd); break;
  }
  IBM_LAUNCH_CHECK();
  return IBM_OK;
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
  // fewer, fatter blocks than forward: every block ends with d*3 atomics
  long long need = ceil_div(M, kThreads / 32);
  long long cap = (long long)sm_count() * 2;
  const int grid = (int)(need < cap ? need : cap);
  const size_t smem = (size_t)3 * 8 * ch * 256 * sizeof(float);
#define IBM_LN_BWD(CHV)                                                                                          \
  do {                                                                                                           \
    static bool attr_set_##CHV = false;                                                                          \
    if (!attr_set_##CHV) {                                                                                       \
      IBM_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel<CHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_set_##CHV = true;                                                                                     \
    }                                                                                                            \
    layernorm_bwd_kernel<CHV><<<grid, kThreads, smem, st>>>(dyp, sp, ld, gamma, mean, rstd, M, d, dsp, dgamma, dbeta, dcolsum); \
  } while (0)
  switch (ch) {
    case 1: IBM_LN_BWD(1); break;
    case 2: IBM_LN_BWD(2); break;
    case 3: IBM_LN_BWD(3); break;
    default: IBM_LN_BWD(4); break;
  }
#undef IBM_LN_BWD
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}
