// Fused regression loss evaluator (forward + backward) — one launch each.
//
// Replaces the ~25 ATen launches + 7 `.item()` syncs per call of
//   /root/reference/src/loss/RegressionLossEvaluator.py:160-221 (step 1: four squared-diff means,
//   the 10 N CoP mask, the component-selected sum) and :230-263 (step 2.2: six last-frame reports).
// HBM-bound: algorithmic traffic 240*F bytes/window forward, +120*F bytes/window for the
// backward write (bf16 grads: +60*F).  Thread mapping: one thread per (row, channel) of the
// 30-channel row [CoP6|F6|tau6|W12], so rows30 layouts are read fully coalesced and any other
// (stride_b, stride_f) layout (FeedForward's quantity-blocked output, separate label tensors) is
// read in 24/48-byte runs.  Column sums: per-thread register accumulator (a thread always owns
// the same channel) → smem → per-block partials → last-arriving block reduces in fp64 in a
// fixed order (deterministic, no float atomics).
#include "common.cuh"

namespace ibm {

struct LossParams {
  const float* out[4];
  const float* lab[4];
  long long osb[4], osf[4], lsb[4], lsf[4];
  float w[30];
  long long B, F;
  float thr;
};

struct GradParams {
  void* g[4];
  long long gsb[4], gsf[4];
};

constexpr int kRowsPerBlock = 8;      // 8 rows x 30 channels = 240 active threads of 256
constexpr int kLossThreads = 256;
constexpr int kResult = 40;           // floats per partial / result

__device__ __forceinline__ void ch_to_qc(int ch, int& q, int& c) {
  if (ch < 6) { q = 0; c = ch; }
  else if (ch < 12) { q = 1; c = ch - 6; }
  else if (ch < 18) { q = 2; c = ch - 12; }
  else { q = 3; c = ch - 18; }
}

// mask_by_threes on the label force (…Evaluator.py:85-108, threshold 10.0 at :205-209): strict >.
__device__ __forceinline__ bool force_mask(const float* f3, float thr) {
  float a = __ldg(f3), b = __ldg(f3 + 1), c = __ldg(f3 + 2);
  float n2 = __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
  return sqrtf(n2) > thr;
}

__global__ void __launch_bounds__(kLossThreads)
loss_fwd_kernel(const LossParams p, float* __restrict__ result, float* __restrict__ partials,
                unsigned int* __restrict__ counter) {
  __shared__ float red[kLossThreads];
  __shared__ float rep_s[kLossThreads / 32][7];
  __shared__ double fin[kResult];
  __shared__ bool is_last;
  const int t = threadIdx.x;
  const long long M = p.B * p.F;

  // ---- phase 1: per-channel squared-error sums over all (b, f) rows -------------------------
  float acc = 0.f;
  if (t < kRowsPerBlock * 30) {
    const int ch = t % 30, r = t / 30;
    int q, c;
    ch_to_qc(ch, q, c);
    const float* __restrict__ ob = p.out[q] + c;
    const float* __restrict__ lb = p.lab[q] + c;
    const float* __restrict__ fb = p.lab[1] + (c / 3) * 3;
    const long long osb = p.osb[q], osf = p.osf[q], lsb = p.lsb[q], lsf = p.lsf[q];
    const long long fsb = p.lsb[1], fsf = p.lsf[1];
    const long long S = (long long)gridDim.x * kRowsPerBlock;
    const long long dS_b = S / p.F, dS_f = S % p.F;
    long long m = (long long)blockIdx.x * kRowsPerBlock + r;
    long long b = m / p.F, f = m % p.F;
#pragma unroll 4
    for (; m < M; m += S) {
      float o = __ldg(ob + b * osb + f * osf);
      float l = __ldg(lb + b * lsb + f * lsf);
      float d = o - l;
      if (q == 0 && !force_mask(fb + b * fsb + f * fsf, p.thr)) d = 0.f;
      acc = fmaf(d, d, acc);
      b += dS_b;
      f += dS_f;
      if (f >= p.F) { f -= p.F; ++b; }
    }
  }
  red[t] = acc;

  // ---- phase 2: last-frame report norms, one thread per window -------------------------------
  float rep[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  {
    const long long f = p.F - 1;
    // windows are dealt round-robin over BLOCKS (thread t of block j takes window j + t*grid), so every SM
    // carries a few of these latency-bound rows next to its streaming work instead of a few blocks carrying all
    for (long long b = (long long)blockIdx.x + (long long)t * gridDim.x; b < p.B; b += (long long)gridDim.x * kLossThreads) {
      float d[30];
      float of[6], lf[6];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int C = q == 3 ? 12 : 6, base = q == 3 ? 18 : 6 * q;
        const float* o = p.out[q] + b * p.osb[q] + f * p.osf[q];
        const float* l = p.lab[q] + b * p.lsb[q] + f * p.lsf[q];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float ov = __ldg(o + c), lv = __ldg(l + c);
          d[base + c] = ov - lv;
          if (q == 1) { of[c] = ov; lf[c] = lv; }
        }
      }
      const float* lforce = p.lab[1] + b * p.lsb[1] + f * p.lsf[1];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (!force_mask(lforce + 3 * g, p.thr)) { d[3 * g] = 0.f; d[3 * g + 1] = 0.f; d[3 * g + 2] = 0.f; }
      }
      auto n3 = [&](int i) { return sqrtf(d[i] * d[i] + d[i + 1] * d[i + 1] + d[i + 2] * d[i + 2]); };
      rep[0] += n3(6) + n3(9);                  // force      (…:232-235)
      rep[1] += n3(12) + n3(15);                // moment     (…:236-239)
      rep[2] += n3(0) + n3(3);                  // masked CoP (…:240-243)
      rep[3] += n3(18);                         // wrench moment, left  (…:244-248)
      rep[4] += n3(24);                         // wrench moment, right (…:249-253)
      float w0 = 0.f, w1 = 0.f;
#pragma unroll
      for (int c = 0; c < 6; ++c) { w0 = fmaf(d[18 + c], d[18 + c], w0); w1 = fmaf(d[24 + c], d[24 + c], w1); }
      rep[5] += sqrtf(w0) + sqrtf(w1);          // wrench v=6 (…:255-259)
      float cx = (of[0] + of[3]) - (lf[0] + lf[3]);
      float cy = (of[1] + of[4]) - (lf[1] + lf[4]);
      float cz = (of[2] + of[5]) - (lf[2] + lf[5]);
      rep[6] += sqrtf(cx * cx + cy * cy + cz * cz);   // CoM acc   (…:143-158, 260-263)
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    float v = warp_sum(rep[k]);
    if ((t & 31) == 0) rep_s[t >> 5][k] = v;
  }
  __syncthreads();

  float* mine = partials + (size_t)blockIdx.x * kResult;
  if (t < 30) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kRowsPerBlock; ++k) s += red[t + 30 * k];
    mine[t] = s;
  } else if (t >= 32 && t < 39) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kLossThreads / 32; ++k) s += rep_s[k][t - 32];
    mine[30 + (t - 32)] = s;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;

  // ---- last block: fixed-order fp64 reduction over all block partials ------------------------
  __threadfence();
  if (t < 37) {
    double s = 0.0;
    for (unsigned int j = 0; j < gridDim.x; ++j) s += (double)__ldcg(partials + (size_t)j * kResult + t);
    fin[t] = s;
  }
  __syncthreads();
  if (t == 0) {
    const double N = (double)M, Bd = (double)p.B;
    double loss = 0.0;
    for (int c = 0; c < 30; ++c) {
      double v = fin[c] / N;
      result[1 + c] = (float)v;
      loss += (double)p.w[c] * v;
    }
    result[0] = (float)loss;
    result[31] = (float)(fin[30] / (2.0 * Bd));                       // force
    result[32] = (float)(fin[31] / (2.0 * Bd));                       // moment
    result[33] = (float)(fin[32] / (2.0 * Bd));                       // cop
    result[34] = (float)(0.5 * (fin[33] / Bd + fin[34] / Bd));        // wrench moment
    result[35] = (float)(fin[35] / (2.0 * Bd));                       // wrench
    result[36] = (float)(fin[36] / Bd);                               // com acc
    result[37] = result[38] = result[39] = 0.f;
    *counter = 0u;
  }
}

template <bool kBf16>
__global__ void __launch_bounds__(kLossThreads)
loss_bwd_kernel(const LossParams p, const GradParams gp, const float* __restrict__ upstream) {
  const int t = threadIdx.x;
  if (t >= kRowsPerBlock * 30) return;
  const long long M = p.B * p.F;
  const int ch = t % 30, r = t / 30;
  int q, c;
  ch_to_qc(ch, q, c);
  const float up = upstream ? __ldg(upstream) : 1.f;
  const float scale = p.w[ch] * up / (float)M;
  const float* __restrict__ ob = p.out[q] + c;
  const float* __restrict__ lb = p.lab[q] + c;
  const float* __restrict__ fb = p.lab[1] + (c / 3) * 3;
  const long long osb = p.osb[q], osf = p.osf[q], lsb = p.lsb[q], lsf = p.lsf[q];
  const long long fsb = p.lsb[1], fsf = p.lsf[1];
  const long long gsb = gp.gsb[q], gsf = gp.gsf[q];
  const long long S = (long long)gridDim.x * kRowsPerBlock;
  const long long dS_b = S / p.F, dS_f = S % p.F;
  long long m = (long long)blockIdx.x * kRowsPerBlock + r;
  long long b = m / p.F, f = m % p.F;
#pragma unroll 4
  for (; m < M; m += S) {
    float d = __ldg(ob + b * osb + f * osf) - __ldg(lb + b * lsb + f * lsf);
    if (q == 0 && !force_mask(fb + b * fsb + f * fsf, p.thr)) d = 0.f;
    float g = (2.f * d) * scale;
    if (kBf16) reinterpret_cast<__nv_bfloat16*>(gp.g[q])[b * gsb + f * gsf + c] = __float2bfloat16_rn(g);
    else reinterpret_cast<float*>(gp.g[q])[b * gsb + f * gsf + c] = g;
    b += dS_b;
    f += dS_f;
    if (f >= p.F) { f -= p.F; ++b; }
  }
}

static int fill_params(LossParams& p, const void* const* h_out, const int64_t* os, const void* const* h_lab,
                       const int64_t* ls, int64_t B, int64_t F, const float* w, float thr) {
  IBM_CHECK_ARG(h_out && os && h_lab && ls && w, "regression_loss: null argument");
  // ValueError cases of the reference helpers (…Evaluator.py:74-79): empty tensors
  IBM_CHECK_ARG(B > 0 && F > 0, "regression_loss: Output and label tensors must not be empty (B=%lld F=%lld)",
                (long long)B, (long long)F);
  for (int q = 0; q < 4; ++q) {
    IBM_CHECK_ARG(h_out[q] && h_lab[q], "regression_loss: null tensor %d", q);
    p.out[q] = static_cast<const float*>(h_out[q]);
    p.lab[q] = static_cast<const float*>(h_lab[q]);
    p.osb[q] = os[2 * q]; p.osf[q] = os[2 * q + 1];
    p.lsb[q] = ls[2 * q]; p.lsf[q] = ls[2 * q + 1];
  }
  for (int c = 0; c < 30; ++c) p.w[c] = w[c];
  p.B = B; p.F = F; p.thr = thr;
  return IBM_OK;
}

static int loss_grid(int64_t B, int64_t F) {
  int64_t need = ceil_div(B * F, kRowsPerBlock);
  int64_t cap = (int64_t)sm_count() * 8;
  int64_t maxp = (int64_t)(ibm_workspace_bytes() - 256) / (kResult * sizeof(float));
  if (cap > maxp) cap = maxp;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace ibm

extern "C" int ibm_regression_loss_fwd(const void* const* h_out, const int64_t* h_out_strides,
                                       const void* const* h_lab, const int64_t* h_lab_strides, int64_t B,
                                       int64_t F, const float* h_weights, float threshold, float* result,
                                       void* workspace, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  LossParams p;
  int rc = fill_params(p, h_out, h_out_strides, h_lab, h_lab_strides, B, F, h_weights, threshold);
  if (rc) return rc;
  IBM_CHECK_ARG(result && workspace, "regression_loss_fwd: null result/workspace");
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  loss_fwd_kernel<<<loss_grid(B, F), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(p, result, partials, counter);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_regression_loss_bwd(const void* const* h_out, const int64_t* h_out_strides,
                                       const void* const* h_lab, const int64_t* h_lab_strides, int64_t B,
                                       int64_t F, const float* h_weights, float threshold, const float* upstream,
                                       void* const* h_grad, const int64_t* h_grad_strides, int32_t grad_dtype,
                                       void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  LossParams p;
  int rc = fill_params(p, h_out, h_out_strides, h_lab, h_lab_strides, B, F, h_weights, threshold);
  if (rc) return rc;
  IBM_CHECK_ARG(h_grad && h_grad_strides, "regression_loss_bwd: null grad");
  IBM_CHECK_ARG(grad_dtype == IBM_F32 || grad_dtype == IBM_BF16, "regression_loss_bwd: bad grad dtype %d", grad_dtype);
  GradParams gp;
  for (int q = 0; q < 4; ++q) {
    IBM_CHECK_ARG(h_grad[q], "regression_loss_bwd: null grad tensor %d", q);
    gp.g[q] = h_grad[q];
    gp.gsb[q] = h_grad_strides[2 * q];
    gp.gsf[q] = h_grad_strides[2 * q + 1];
  }
  int grid = loss_grid(B, F);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (grad_dtype == IBM_BF16) loss_bwd_kernel<true><<<grid, kLossThreads, 0, s>>>(p, gp, upstream);
  else loss_bwd_kernel<false><<<grid, kLossThreads, 0, s>>>(p, gp, upstream);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}
