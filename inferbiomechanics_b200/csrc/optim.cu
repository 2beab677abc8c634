// Fused optimizer step over the flat fp32 parameter arena.  One launch updates parameters and
// optimizer state in place, applies the data-parallel 1/world_size gradient scale, and refreshes
// the bf16 shadow weights the tcgen05 GEMMs read.  HBM-bound (RMSprop: 12 B read + 10 B written
// per parameter).
//
// Replaces torch.optim.{RMSprop,Adam,SGD,Adagrad,Adadelta,Adamax}(lr) with torch defaults as
// constructed at /root/reference/src/cli/train.py:183-197 and stepped at :284.
#include "common.cuh"

namespace ibm {

constexpr int kThreads = 256;

struct OptArgs {
  float lr, gscale;
  float bc1, bc2_sqrt;   // Adam / Adamax bias corrections for this step
  const long long* step_dev;   // optional device-resident step count (CUDA-graph replays): the corrections are derived from it
};

template <int KIND>
__device__ __forceinline__ void opt_update(float& p, float g, float& s0, float& s1, const OptArgs& a) {
  if (KIND == 0) {            // RMSprop(alpha=.99, eps=1e-8)
    s0 = fmaf(0.99f, s0, 0.01f * g * g);               // 1-alpha evaluated in fp32 like torch's python float → 0.010000000000000009
    p -= a.lr * (g / (sqrtf(s0) + 1e-8f));
  } else if (KIND == 1) {     // Adam(.9, .999, 1e-8)
    s0 = s0 + (1.f - 0.9f) * (g - s0);
    s1 = fmaf(0.999f, s1, (1.f - 0.999f) * g * g);
    const float denom = sqrtf(s1) / a.bc2_sqrt + 1e-8f;
    p -= (a.lr / a.bc1) * (s0 / denom);
  } else if (KIND == 2) {     // SGD
    p -= a.lr * g;
  } else if (KIND == 3) {     // Adagrad(eps=1e-10)
    s0 = fmaf(g, g, s0);
    p -= a.lr * (g / (sqrtf(s0) + 1e-10f));
  } else if (KIND == 4) {     // Adadelta(rho=.9, eps=1e-6)
    s0 = fmaf(0.9f, s0, (1.f - 0.9f) * g * g);
    const float delta = sqrtf(s1 + 1e-6f) / sqrtf(s0 + 1e-6f) * g;
    s1 = fmaf(0.9f, s1, (1.f - 0.9f) * delta * delta);
    p -= a.lr * delta;
  } else {                    // Adamax(.9, .999, 1e-8)
    s0 = s0 + (1.f - 0.9f) * (g - s0);
    s1 = fmaxf(0.999f * s1, fabsf(g) + 1e-8f);
    p -= (a.lr / a.bc1) * (s0 / s1);
  }
}

template <int KIND>
__global__ void __launch_bounds__(kThreads)
optimizer_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ st0,
                 float* __restrict__ st1, __nv_bfloat16* __restrict__ pb, long long n, OptArgs a) {
  if constexpr (KIND == 1 || KIND == 5) {
    if (a.step_dev != nullptr) {
      // one thread per block evaluates the two powers in double precision (as the host does for a by-value step)
      __shared__ float s_bc[2];
      if (threadIdx.x == 0) {
        const double st = (double)__ldg(a.step_dev);
        s_bc[0] = (float)(1.0 - pow(0.9, st));
        s_bc[1] = (float)sqrt(1.0 - pow(0.999, st));
      }
      __syncthreads();
      a.bc1 = s_bc[0];
      a.bc2_sqrt = s_bc[1];
    }
  }
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  constexpr bool kS0 = KIND != 2, kS1 = (KIND == 1 || KIND == 4 || KIND == 5);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = *reinterpret_cast<const float4*>(param + 4 * i);
    float4 g = ld_stream_f4(grad + 4 * i);
    float4 s0 = kS0 ? *reinterpret_cast<const float4*>(st0 + 4 * i) : make_float4(0, 0, 0, 0);
    float4 s1 = kS1 ? *reinterpret_cast<const float4*>(st1 + 4 * i) : make_float4(0, 0, 0, 0);
    opt_update<KIND>(p.x, g.x * a.gscale, s0.x, s1.x, a);
    opt_update<KIND>(p.y, g.y * a.gscale, s0.y, s1.y, a);
    opt_update<KIND>(p.z, g.z * a.gscale, s0.z, s1.z, a);
    opt_update<KIND>(p.w, g.w * a.gscale, s0.w, s1.w, a);
    *reinterpret_cast<float4*>(param + 4 * i) = p;
    if (kS0) *reinterpret_cast<float4*>(st0 + 4 * i) = s0;
    if (kS1) *reinterpret_cast<float4*>(st1 + 4 * i) = s1;
    if (pb) *reinterpret_cast<uint2*>(pb + 4 * i) = make_uint2(pack_bf16x2(p.x, p.y), pack_bf16x2(p.z, p.w));
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float p = param[i], s0 = kS0 ? st0[i] : 0.f, s1 = kS1 ? st1[i] : 0.f;
    opt_update<KIND>(p, grad[i] * a.gscale, s0, s1, a);
    param[i] = p;
    if (kS0) st0[i] = s0;
    if (kS1) st1[i] = s1;
    if (pb) pb[i] = __float2bfloat16_rn(p);
  }
}

}  // namespace ibm

static int optimizer_step_impl(int32_t kind, float* param, const float* grad, float* state0, float* state1, void* param_bf16,
                               int64_t n, float lr, float grad_scale, int64_t step, const int64_t* step_dev, void* stream);

extern "C" int ibm_optimizer_step(int32_t kind, float* param, const float* grad, float* state0, float* state1,
                                  void* param_bf16, int64_t n, float lr, float grad_scale, int64_t step, void* stream) {
  return optimizer_step_impl(kind, param, grad, state0, state1, param_bf16, n, lr, grad_scale, step, nullptr, stream);
}

extern "C" int ibm_optimizer_step_dev(int32_t kind, float* param, const float* grad, float* state0, float* state1,
                                      void* param_bf16, int64_t n, float lr, float grad_scale, const int64_t* step_dev, void* stream) {
  if (step_dev == nullptr) {
    ibm::set_error("optimizer_step_dev: null step counter");
    return IBM_E_ARG;
  }
  return optimizer_step_impl(kind, param, grad, state0, state1, param_bf16, n, lr, grad_scale, 1, step_dev, stream);
}

static int optimizer_step_impl(int32_t kind, float* param, const float* grad, float* state0, float* state1, void* param_bf16,
                               int64_t n, float lr, float grad_scale, int64_t step, const int64_t* step_dev, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(kind >= 0 && kind <= 5, "optimizer_step: unknown optimizer kind %d", kind);
  IBM_CHECK_ARG(param && grad && n > 0 && step >= 1, "optimizer_step: bad argument");
  IBM_CHECK_ARG(kind == 2 || state0, "optimizer_step: state0 required");
  IBM_CHECK_ARG(!(kind == 1 || kind == 4 || kind == 5) || state1, "optimizer_step: state1 required");
  IBM_CHECK_ARG(aligned16(param) && aligned16(grad) && (!state0 || aligned16(state0)) && (!state1 || aligned16(state1)) &&
                    (!param_bf16 || (reinterpret_cast<uintptr_t>(param_bf16) % 8 == 0)),
                "optimizer_step: arenas must be 16-byte aligned");
  OptArgs a;
  a.lr = lr;
  a.gscale = grad_scale;
  a.bc1 = (float)(1.0 - pow(0.9, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(0.999, (double)step));
  a.step_dev = reinterpret_cast<const long long*>(step_dev);
  long long need = ceil_div(n / 4 + 1, kThreads);
  long long cap = (long long)sm_count() * 16;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto* pb = static_cast<__nv_bfloat16*>(param_bf16);
  switch (kind) {
    case 0: optimizer_kernel<0><<<grid, kThreads, 0, s>>>(param, grad, state0, state1, pb, n, a); break;
    case 1: optimizer_kernel<1><<<grid, kThreads, 0, s>>>(param, grad, state0, state1, pb, n, a); break;
    case 2: optimizer_kernel<2><<<grid, kThreads, 0, s>>>(param, grad, state0, state1, pb, n, a); break;
    case 3: optimizer_kernel<3><<<grid, kThreads, 0, s>>>(param, grad, state0, state1, pb, n, a); break;
    case 4: optimizer_kernel<4><<<grid, kThreads, 0, s>>>(param, grad, state0, state1, pb, n, a); break;
    default: optimizer_kernel<5><<<grid, kThreads, 0, s>>>(param, grad, state0, state1, pb, n, a); break;
  }
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}
