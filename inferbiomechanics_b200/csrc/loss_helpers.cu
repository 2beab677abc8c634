// The four static helpers of the reference loss evaluator with their GENERAL contract
// (/root/reference/src/loss/RegressionLossEvaluator.py:73-158): any channel count C for the squared-diff
// mean, any C % 3 == 0 for the mask, any C % vec_size == 0 for the last-frame mean norm.  The fused
// kernels of loss.cu serve RegressionLossEvaluator.__call__ (fixed 6/6/6/12 channels); these serve the
// helpers called on their own — the shapes of the reference's unit tests
// (test/loss/test_RegressionLossEvaluator.py: (2,4,3), (1,2,3), (1,1,6) …) go through here.
// All HBM-bound streaming reductions: a block strides over rows, a thread owns one channel (coalesced
// along the channel-contiguous rows), block partials go to the workspace and the last-arriving block
// reduces them in fp64 in a fixed order (deterministic; no float atomics) — the scheme of loss.cu.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace ibm {

constexpr int kHelperThreads = 256;

struct View3 {        // (B, F, C) fp32 view with unit channel stride
  const float* p;
  long long sb, sf;
};

__device__ __forceinline__ bool last_block_arrives(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// ---- get_squared_diff_mean_vector (…Evaluator.py:73-83): out[c] = mean_{b,f} (o-l)^2 -------------------------
// grid = (row chunks, channel tiles of 32); block = 32 channels x 8 row lanes.
__global__ void __launch_bounds__(kHelperThreads)
sqdiff_mean_kernel(View3 o, View3 l, long long B, long long F, int C, float* __restrict__ out,
                   float* __restrict__ partials, unsigned int* __restrict__ counter) {
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + cx;
  const long long rows = B * F;
  float acc = 0.f;
  if (c < C) {
    for (long long r = (long long)blockIdx.x * 8 + ry; r < rows; r += (long long)gridDim.x * 8) {
      const long long b = r / F, f = r - b * F;
      const float d = __ldg(o.p + b * o.sb + f * o.sf + c) - __ldg(l.p + b * l.sb + f * l.sf + c);
      acc = fmaf(d, d, acc);
    }
  }
  __shared__ float sm[8][33];
  sm[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][cx];
    partials[(size_t)blockIdx.x * C + c] = s;
  }
  if (!last_block_arrives(counter)) return;
  for (int cc = threadIdx.x; cc < C; cc += kHelperThreads) {
    double s = 0.0;
    for (unsigned int j = 0; j < gridDim.x; ++j) s += (double)__ldcg(partials + (size_t)j * C + cc);
    out[cc] = (float)(s / (double)rows);
  }
  if (threadIdx.x == 0) *counter = 0u;
}

// d/d o of  sum_c up[c] * mean_{b,f}(o-l)^2  =  2 up[c] (o-l) / (B F);  d/d l is its negative
__global__ void __launch_bounds__(kHelperThreads)
sqdiff_mean_bwd_kernel(View3 o, View3 l, long long B, long long F, int C, const float* __restrict__ up,
                       float* __restrict__ go, float* __restrict__ gl) {
  const long long n = B * F * C;
  const float inv = 2.f / (float)(B * F);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const long long b = r / F, f = r - b * F;
    const float d = __ldg(o.p + b * o.sb + f * o.sf + c) - __ldg(l.p + b * l.sb + f * l.sf + c);
    const float g = __ldg(up + c) * inv * d;
    if (go) go[i] = g;
    if (gl) gl[i] = -g;
  }
}

// ---- get_mask_by_threes (…Evaluator.py:85-108): mask[b,f,3g:3g+3] = (||x[b,f,3g:3g+3]||_2 > thr) ---------------
// strict >, norm accumulated as (a*a + b*b) + c*c without FMA contraction, like loss.cu's CoP mask
__global__ void __launch_bounds__(kHelperThreads)
mask_by_threes_kernel(View3 x, long long B, long long F, int G, float thr, float* __restrict__ out) {
  const long long n = B * F * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / G;
    const int g = (int)(i - r * G);
    const long long b = r / F, f = r - b * F;
    const float* p = x.p + b * x.sb + f * x.sf + 3 * g;
    const float a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2));
    const float m = sqrtf(n2) > thr ? 1.f : 0.f;
    float* q = out + r * (3ll * G) + 3 * g;
    q[0] = m; q[1] = m; q[2] = m;
  }
}

// ---- get_mean_norm_error (…Evaluator.py:119-141) and get_com_acc_error (:143-158) ---------------------------
// mean over (b, g) of || (o-l)[b, F-1, g*v:(g+1)*v] ||_2 — the LAST frame only (:136).
// fold: C == 6, d[k] = (o[k]+o[k+3]) - (l[k]+l[k+3]), one 3-vector per window (:154-158).
__global__ void __launch_bounds__(kHelperThreads)
mean_norm_kernel(View3 o, View3 l, long long B, long long F, int C, int v, int fold, float* __restrict__ out,
                 float* __restrict__ partials, unsigned int* __restrict__ counter) {
  const int G = fold ? 1 : C / v;
  const long long n = B * G;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / G;
    const int g = (int)(i - b * G);
    const float* po = o.p + b * o.sb + (F - 1) * o.sf + g * v;
    const float* pl = l.p + b * l.sb + (F - 1) * l.sf + g * v;
    float s = 0.f;
    if (fold) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float d = (__ldg(po + k) + __ldg(po + k + 3)) - (__ldg(pl + k) + __ldg(pl + k + 3));
        s = fmaf(d, d, s);
      }
    } else {
      for (int k = 0; k < v; ++k) {
        const float d = __ldg(po + k) - __ldg(pl + k);
        s = fmaf(d, d, s);
      }
    }
    acc += sqrtf(s);
  }
  acc = warp_sum(acc);
  __shared__ float sm[kHelperThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kHelperThreads / 32; ++k) s += sm[k];
    partials[blockIdx.x] = s;
  }
  if (!last_block_arrives(counter)) return;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (unsigned int j = 0; j < gridDim.x; ++j) s += (double)__ldcg(partials + j);
    out[0] = (float)(s / (double)n);
    *counter = 0u;
  }
}

static int helper_grid(int64_t items, int64_t per_block, int64_t floats_per_partial) {
  int64_t need = ceil_div(items, per_block);
  int64_t cap = (int64_t)sm_count() * 8;
  int64_t maxp = (int64_t)(ibm_workspace_bytes() - 256) / (int64_t)(floats_per_partial * sizeof(float));
  if (cap > maxp) cap = maxp;
  if (cap < 1) cap = 1;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace ibm

extern "C" int ibm_sqdiff_mean_vector(const float* out_t, int64_t o_sb, int64_t o_sf, const float* lab_t, int64_t l_sb,
                                      int64_t l_sf, int64_t B, int64_t F, int32_t C, float* result, void* workspace,
                                      void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(out_t && lab_t && result && workspace, "sqdiff_mean_vector: null argument");
  IBM_CHECK_ARG(B > 0 && F > 0 && C > 0, "Output and label tensors must not be empty");
  IBM_CHECK_ARG((int64_t)C * sizeof(float) <= ibm_workspace_bytes() - 256, "sqdiff_mean_vector: C=%d exceeds the workspace", C);
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  dim3 grid(helper_grid(B * F, 64, C), (unsigned)ceil_div(C, 32));
  sqdiff_mean_kernel<<<grid, kHelperThreads, 0, static_cast<cudaStream_t>(stream)>>>(View3{out_t, o_sb, o_sf}, View3{lab_t, l_sb, l_sf},
                                                                                     B, F, C, result, partials, counter);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_sqdiff_mean_vector_bwd(const float* out_t, int64_t o_sb, int64_t o_sf, const float* lab_t, int64_t l_sb,
                                          int64_t l_sf, int64_t B, int64_t F, int32_t C, const float* upstream,
                                          float* grad_out, float* grad_lab, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(out_t && lab_t && upstream && (grad_out || grad_lab), "sqdiff_mean_vector_bwd: null argument");
  IBM_CHECK_ARG(B > 0 && F > 0 && C > 0, "Output and label tensors must not be empty");
  const int grid = (int)std::min<int64_t>(ceil_div(B * F * C, kHelperThreads), (int64_t)sm_count() * 8);
  sqdiff_mean_bwd_kernel<<<grid, kHelperThreads, 0, static_cast<cudaStream_t>(stream)>>>(View3{out_t, o_sb, o_sf}, View3{lab_t, l_sb, l_sf},
                                                                                         B, F, C, upstream, grad_out, grad_lab);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_mask_by_threes(const float* x, int64_t sb, int64_t sf, int64_t B, int64_t F, int32_t C, float threshold,
                                  float* mask, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(x && mask, "mask_by_threes: null argument");
  IBM_CHECK_ARG(B > 0 && F > 0 && C > 0, "Mask tensor must not be empty");
  IBM_CHECK_ARG(C % 3 == 0, "Mask tensor must have a final dimension divisible by 3");
  const int G = C / 3;
  const int grid = (int)std::min<int64_t>(ceil_div(B * F * G, kHelperThreads), (int64_t)sm_count() * 8);
  mask_by_threes_kernel<<<grid, kHelperThreads, 0, static_cast<cudaStream_t>(stream)>>>(View3{x, sb, sf}, B, F, G, threshold, mask);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_mean_norm_error(const float* out_t, int64_t o_sb, int64_t o_sf, const float* lab_t, int64_t l_sb, int64_t l_sf,
                                   int64_t B, int64_t F, int32_t C, int32_t vec_size, int32_t fold_halves, float* result,
                                   void* workspace, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(out_t && lab_t && result && workspace, "mean_norm_error: null argument");
  IBM_CHECK_ARG(B > 0 && F > 0 && C > 0, "Output and label tensors must not be empty");
  if (fold_halves) IBM_CHECK_ARG(C == 6, "Output and label tensors must have a 6 dimensional final dimension");
  else IBM_CHECK_ARG(vec_size > 0 && C % vec_size == 0, "Tensors must have a final dimension divisible by vec_size=%d", vec_size);
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  float* partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  const int64_t n = B * (fold_halves ? 1 : C / vec_size);
  const int grid = helper_grid(n, kHelperThreads, 1);
  mean_norm_kernel<<<grid, kHelperThreads, 0, static_cast<cudaStream_t>(stream)>>>(View3{out_t, o_sb, o_sf}, View3{lab_t, l_sb, l_sf}, B, F,
                                                                                   C, vec_size, fold_halves, result, partials, counter);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}
