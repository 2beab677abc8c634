// BatchNorm1d over packed bf16 rows (nn.BatchNorm1d on the layer input, FeedForwardRegressionBaseline.py:71-72).
//
// Column statistics of a row-major [M, ld] bf16 matrix: the rows are cut into R chunks, every (64-column, chunk) block
// makes two passes over its chunk (mean, then centred sum of squares — the chunk stays in L2) and the chunks are merged
// with Chan's parallel-variance update, so the result does not suffer the E[x^2]-E[x]^2 cancellation.  Each lane owns a
// bf16x2 column pair: a warp reads one 128-byte row segment per instruction.
#include "common.cuh"

namespace ibm {
namespace {

constexpr int kWarps = 8;
constexpr int kColsPerBlock = 64;
constexpr int kMaxChunks = 64;

__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16* p) {
  return unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p));
}

// sums over the kWarps row-lanes of a block: s[w][lane] -> total in every thread of the column pair
__device__ __forceinline__ float2 block_colsum(float2 v, float2 (*sm)[32]) {
  sm[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  float2 t = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    t.x += sm[w][threadIdx.x].x;
    t.y += sm[w][threadIdx.x].y;
  }
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(32 * kWarps)
bn_chunk_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long M, int C, int rows_per_chunk,
                      float* __restrict__ ws /* [chunks][2][Cp] */, int Cp) {
  __shared__ float2 sm[kWarps][32];
  const int c = blockIdx.x * kColsPerBlock + 2 * threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(M, r0 + rows_per_chunk);
  const bool live = c < C;                         // ld is a multiple of 8, so the pair (c, c+1) is always inside the row
  float2 s = make_float2(0.f, 0.f);
  if (live)
    for (long long r = r0 + threadIdx.y; r < r1; r += kWarps) {
      const float2 v = ld_pair(x + r * ld + c);
      s.x += v.x;
      s.y += v.y;
    }
  s = block_colsum(s, sm);
  const float inv = 1.f / (float)(r1 - r0);
  const float2 mean = make_float2(s.x * inv, s.y * inv);
  float2 q = make_float2(0.f, 0.f);
  if (live)
    for (long long r = r0 + threadIdx.y; r < r1; r += kWarps) {
      const float2 v = ld_pair(x + r * ld + c);
      q.x += (v.x - mean.x) * (v.x - mean.x);
      q.y += (v.y - mean.y) * (v.y - mean.y);
    }
  q = block_colsum(q, sm);
  if (live && threadIdx.y == 0) {
    float* w = ws + (size_t)blockIdx.y * 2 * Cp;
    w[c] = mean.x;
    w[c + 1] = mean.y;
    w[Cp + c] = q.x;
    w[Cp + c + 1] = q.y;
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ ws, int Cp, int C, long long M, int rows_per_chunk, int chunks,
                                   float momentum, float eps, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int k = 0; k < chunks; ++k) {
    const float nb = (float)min((long long)rows_per_chunk, M - (long long)k * rows_per_chunk);
    const float mb = ws[(size_t)k * 2 * Cp + c], qb = ws[(size_t)k * 2 * Cp + Cp + c];
    const float tot = n + nb, delta = mb - mean;
    mean += delta * (nb / tot);
    m2 += qb + delta * delta * (n * nb / tot);
    n = tot;
  }
  save_mean[c] = mean;
  save_rstd[c] = rsqrtf(m2 / n + eps);                      // biased variance normalises (torch semantics)
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (m2 / (n - 1.f));   // unbiased
}

// y = (x - mean) * rstd * gamma + beta;  eval mode: mean/var are the running statistics
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy, long long M,
                int C, int pairs, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ mean, const float* __restrict__ rstd_or_var, int var_given, float eps) {
  const long long total = M * pairs;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / pairs;
    const int c = 2 * (int)(i - r * pairs);
    float2 v = make_float2(0.f, 0.f);
    if (c < C) v = ld_pair(x + r * ldx + c);       // pairs beyond C only zero the pad columns of y
    float o[2] = {0.f, 0.f};
    const float in[2] = {v.x, v.y};
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (c + j < C) {
        const float rs = var_given ? rsqrtf(__ldg(rstd_or_var + c + j) + eps) : __ldg(rstd_or_var + c + j);
        o[j] = (in[j] - __ldg(mean + c + j)) * rs * __ldg(gamma + c + j) + __ldg(beta + c + j);
      }
    *reinterpret_cast<uint32_t*>(y + r * ldy + c) = pack_bf16x2(o[0], o[1]);
  }
}

// s1[c] = sum_m dy, s2[c] = sum_m dy * xhat   (atomics into a zeroed workspace)
__global__ void __launch_bounds__(32 * kWarps)
bn_bwd_sums_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x, long long ldx,
                   long long M, int C, int rows_per_chunk, const float* __restrict__ mean,
                   const float* __restrict__ rstd_or_var, int var_given, float eps, float* __restrict__ ws, int Cp) {
  __shared__ float2 sm[kWarps][32];
  const int c = blockIdx.x * kColsPerBlock + 2 * threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(M, r0 + rows_per_chunk);
  const bool live = c < C;
  float mu[2] = {0.f, 0.f}, rs[2] = {0.f, 0.f};
  if (live)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (c + j < C) {
        mu[j] = mean[c + j];
        rs[j] = var_given ? rsqrtf(rstd_or_var[c + j] + eps) : rstd_or_var[c + j];
      }
  float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
  if (live)
    for (long long r = r0 + threadIdx.y; r < r1; r += kWarps) {
      const float2 g = ld_pair(dy + r * lddy + c), v = ld_pair(x + r * ldx + c);
      s1.x += g.x;
      s1.y += g.y;
      s2.x += g.x * (v.x - mu[0]) * rs[0];
      s2.y += g.y * (v.y - mu[1]) * rs[1];
    }
  s1 = block_colsum(s1, sm);
  s2 = block_colsum(s2, sm);
  if (live && threadIdx.y == 0) {
    atomicAdd(ws + c, s1.x);
    atomicAdd(ws + Cp + c, s2.x);
    if (c + 1 < C) {
      atomicAdd(ws + c + 1, s1.y);
      atomicAdd(ws + Cp + c + 1, s2.y);
    }
  }
}

// dx = gamma * rstd * (dy - s1/M - xhat * s2/M)  [training]   |   gamma * rstd * dy  [eval];  optionally * act'(act_out).
// Same (64 columns x row chunk) decomposition as the statistics kernels, so the column sums of the fp32 dx — the bias
// gradient of the Linear whose activation feeds this BatchNorm — are taken BEFORE the bf16 rounding: in training mode
// sum_m dx is a sum of cancelling terms (sum_m of the BatchNorm input gradient is 0 by construction) and rounding each
// term to bf16 first would leave mostly rounding noise.
__global__ void __launch_bounds__(32 * kWarps)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ x, long long ldx,
                    __nv_bfloat16* __restrict__ dx, long long lddx, long long M, int C, int rows_per_chunk,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd_or_var,
                    int var_given, float eps, const float* __restrict__ ws, int Cp, int training,
                    const __nv_bfloat16* __restrict__ act_out, long long ldact, int act, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, float* __restrict__ dx_colsum) {
  __shared__ float2 sm[kWarps][32];
  const int c = blockIdx.x * kColsPerBlock + 2 * threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = min(M, r0 + rows_per_chunk);
  const float invM = 1.f / (float)M;
  float mu[2] = {0.f, 0.f}, rs[2] = {0.f, 0.f}, ga[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 2; ++j)
    if (c + j < C) {
      mu[j] = mean[c + j];
      rs[j] = var_given ? rsqrtf(rstd_or_var[c + j] + eps) : rstd_or_var[c + j];
      ga[j] = gamma[c + j] * rs[j];
      s1[j] = ws[c + j] * invM;
      s2[j] = ws[Cp + c + j] * invM;
      if (blockIdx.y == 0 && threadIdx.y == 0) {     // parameter gradients accumulate (+=) like every other grad slot
        if (dbeta) dbeta[c + j] += ws[c + j];
        if (dgamma) dgamma[c + j] += ws[Cp + c + j];
      }
    }
  if (!dx) return;
  float2 cs = make_float2(0.f, 0.f);
  if (c < lddx)                                        // pairs in [C, lddx) only zero the pad columns of dx
    for (long long r = r0 + threadIdx.y; r < r1; r += kWarps) {
      float o[2] = {0.f, 0.f};
      if (c < C) {
        const float2 g = ld_pair(dy + r * lddy + c), v = ld_pair(x + r * ldx + c);
        float2 a = make_float2(1.f, 1.f);
        if (act_out) a = ld_pair(act_out + r * ldact + c);
        const float gi[2] = {g.x, g.y}, vi[2] = {v.x, v.y}, ai[2] = {a.x, a.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float d = gi[j];
          if (training) d -= s1[j] + (vi[j] - mu[j]) * rs[j] * s2[j];
          d *= ga[j];                                  // ga == 0 for the odd tail column c + 1 == C
          if (act_out) d *= act_grad_from_output(ai[j], act);
          o[j] = d;
        }
        cs.x += o[0];
        cs.y += o[1];
      }
      *reinterpret_cast<uint32_t*>(dx + r * lddx + c) = pack_bf16x2(o[0], o[1]);
    }
  if (dx_colsum) {
    cs = block_colsum(cs, sm);
    if (threadIdx.y == 0 && c < C) {
      atomicAdd(dx_colsum + c, cs.x);
      if (c + 1 < C) atomicAdd(dx_colsum + c + 1, cs.y);
    }
  }
}

inline int chunks_for(long long M, int* rows_per_chunk) {
  long long rpc = 256;
  if (ceil_div(M, rpc) > kMaxChunks) rpc = ceil_div(ceil_div(M, (long long)kMaxChunks), 8) * 8;
  *rows_per_chunk = (int)rpc;
  return (int)ceil_div(M, rpc);
}

inline int ew_blocks(long long n) {
  long long b = ceil_div(n, 256);
  const long long cap = (long long)sm_count() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace ibm

extern "C" size_t ibm_batchnorm_workspace_floats(int32_t C) {
  return (size_t)2 * (size_t)((C + 1) & ~1) * ibm::kMaxChunks;
}

extern "C" int ibm_batchnorm_fwd(const void* x, int64_t ldx, void* y, int64_t ldy, int64_t M, int32_t C, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var, float* save_mean,
                                 float* save_rstd, int32_t training, float momentum, float eps, float* workspace,
                                 void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(x && y && gamma && beta && M > 0 && C > 0, "batchnorm_fwd: null pointer or empty shape");
  IBM_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "batchnorm_fwd: row strides must be multiples of 8 and >= C");
  auto st = static_cast<cudaStream_t>(stream);
  const int Cp = (C + 1) & ~1;
  const int pairs = (int)(ldy / 2);                 // the whole output row: pad columns are written 0
  if (training) {
    // torch raises "Expected more than 1 value per channel when training" (ValueError) for a single row
    IBM_CHECK_ARG(M > 1, "batchnorm_fwd: Expected more than 1 value per channel when training, got M = 1");
    IBM_CHECK_ARG(save_mean && save_rstd && workspace, "batchnorm_fwd: training needs save_mean, save_rstd and a workspace");
    int rpc;
    const int chunks = chunks_for(M, &rpc);
    dim3 grid((unsigned)ceil_div(C, kColsPerBlock), (unsigned)chunks), block(32, kWarps);
    bn_chunk_stats_kernel<<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, M, C, rpc, workspace, Cp);
    IBM_LAUNCH_CHECK();
    bn_finalize_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(workspace, Cp, C, M, rpc, chunks, momentum, eps, running_mean,
                                                                    running_var, save_mean, save_rstd);
    IBM_LAUNCH_CHECK();
    bn_apply_kernel<<<ew_blocks(M * pairs), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y),
                                                           ldy, M, C, pairs, gamma, beta, save_mean, save_rstd, 0, eps);
  } else {
    IBM_CHECK_ARG(running_mean && running_var, "batchnorm_fwd: eval mode needs the running statistics");
    bn_apply_kernel<<<ew_blocks(M * pairs), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y),
                                                           ldy, M, C, pairs, gamma, beta, running_mean, running_var, 1, eps);
  }
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_batchnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx, void* dx, int64_t lddx, int64_t M,
                                 int32_t C, const float* gamma, const float* mean, const float* rstd_or_var, int32_t training,
                                 float eps, const void* act_out, int64_t ldact, int32_t act, float* dgamma, float* dbeta,
                                 float* dx_colsum, float* workspace, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(dy && x && gamma && mean && rstd_or_var && workspace && M > 0 && C > 0, "batchnorm_bwd: null pointer or empty shape");
  IBM_CHECK_ARG(lddy % 8 == 0 && ldx % 8 == 0 && lddy >= C && ldx >= C && (!dx || (lddx % 8 == 0 && lddx >= C)),
                "batchnorm_bwd: row strides must be multiples of 8 and >= C");
  IBM_CHECK_ARG(!act_out || (ldact % 8 == 0 && ldact >= C && act >= 0 && act <= IBM_ACT_ELU),
                "batchnorm_bwd: activation output needs a stride multiple of 8 and an activation with an output-form derivative");
  auto st = static_cast<cudaStream_t>(stream);
  const int Cp = (C + 1) & ~1;
  const int var_given = training ? 0 : 1;          // eval: (running_mean, running_var); training: (save_mean, save_rstd)
  IBM_CHECK_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * 2 * Cp, st));
  int rpc;
  const int chunks = chunks_for(M, &rpc);
  dim3 grid((unsigned)ceil_div(C, kColsPerBlock), (unsigned)chunks), block(32, kWarps);
  bn_bwd_sums_kernel<<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), lddy, static_cast<const __nv_bfloat16*>(x), ldx,
                                             M, C, rpc, mean, rstd_or_var, var_given, eps, workspace, Cp);
  IBM_LAUNCH_CHECK();
  dim3 grid2((unsigned)ceil_div(dx ? lddx : C, kColsPerBlock), dx ? (unsigned)chunks : 1u);
  bn_bwd_apply_kernel<<<grid2, block, 0, st>>>(
      static_cast<const __nv_bfloat16*>(dy), lddy, static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(dx), lddx,
      M, C, rpc, gamma, mean, rstd_or_var, var_given, eps, workspace, Cp, training, static_cast<const __nv_bfloat16*>(act_out),
      ldact, act, dgamma, dbeta, dx ? dx_colsum : nullptr);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}
