// Window batcher kernels: candidate-window validity, frame gather into packed model inputs, and
// label rows.  Bit-exact integer/byte/copy work, HBM-bound.
//
// Replaces, for data already resident in an HBM frame store:
//   /root/reference/src/data/AddBiomechanicsDataset.py:131-139  (window enumeration predicate)
//   /root/reference/src/data/AddBiomechanicsDataset.py:161-285  (__getitem__: ~18 row_stacks/window)
//   /root/reference/src/models/FeedForwardRegressionBaseline.py:97-108 (concat + reshape + H2D)
#include "common.cuh"

namespace ibm {

constexpr int kThreads = 256;

// valid[c] = !any(missing[ws : ws+T : s])  — python slice semantics: indices ws, ws+s, … < ws+T.
// The caller only enumerates candidates with ws + T < L (Dataset.py:135), so no bound check on L.
__global__ void __launch_bounds__(kThreads)
window_valid_kernel(const uint8_t* __restrict__ missing, const long long* __restrict__ trial_base,
                    const int32_t* __restrict__ cand_trial, const int32_t* __restrict__ cand_start, long long n,
                    int T, int s, uint8_t* __restrict__ valid) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint8_t* m = missing + __ldg(trial_base + __ldg(cand_trial + i)) + __ldg(cand_start + i);
    uint8_t any = 0;
    for (int k = 0; k < T; k += s) any |= __ldg(m + k);
    valid[i] = any ? 0 : 1;
  }
}

// One warp per (window, frame) source row.  Source rows are `stride` frames apart; each row is
// read as float4 (16-byte vectorised; frame_ld % 4 == 0) and written once as fp32 and/or bf16.
template <bool kVec4>
__global__ void __launch_bounds__(kThreads)
pack_windows_kernel(const float* __restrict__ frames, long long frame_ld, int C, const long long* __restrict__ row0,
                    long long n_rows, int F, int stride, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ out_bf16, long long fs, long long we, long long col0) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n_rows; r += warps) {
    const long long i = r / F;
    const int f = (int)(r - i * F);
    const float* src = frames + (__ldg(row0 + i) + (long long)f * stride) * frame_ld;
    float* d32 = out_f32 ? out_f32 + r * C : nullptr;
    __nv_bfloat16* d16 = out_bf16 ? out_bf16 + r * fs + i * we + col0 : nullptr;
    if (kVec4) {
      // source rows 16-byte aligned; fp32 dst (if any) 16-byte aligned per row; bf16 dst 4-byte aligned
      for (int c = lane * 4; c < C; c += 128) {
        if (c + 4 <= C) {
          float4 v = ld_stream_f4(src + c);
          if (d32) st_stream_f4(d32 + c, v);
          if (d16) {
            *reinterpret_cast<uint32_t*>(d16 + c) = pack_bf16x2(v.x, v.y);
            *reinterpret_cast<uint32_t*>(d16 + c + 2) = pack_bf16x2(v.z, v.w);
          }
        } else {                                   // ragged row tail (C % 4 != 0)
          for (int k = c; k < C; ++k) {
            float v = __ldg(src + k);
            if (d32) d32[k] = v;
            if (d16) d16[k] = __float2bfloat16_rn(v);
          }
        }
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        float v = __ldg(src + c);
        if (d32) d32[c] = v;
        if (d16) d16[c] = __float2bfloat16_rn(v);
      }
    }
  }
}

// Label rows30: one WARP per output row (lane = channel, nb = 2 -> 30 of 32 lanes; more bodies loop): the row index, the
// window's frame-store row, its contact-body map and its mass are computed / fetched once per warp instead of once per
// element (the thread-per-element version spent its time in 64-bit divisions: 71 us for 49 MB, 10 % of the copy roofline),
// the 120-byte raw row and the 120-byte output row are each one coalesced access.
// Raw per-frame layout: [cop 3nb | force 3nb | torque 3nb | wrench 6nb].
__global__ void __launch_bounds__(kThreads)
pack_labels_kernel(const float* __restrict__ raw, long long raw_ld, int nb, const long long* __restrict__ row0,
                   const int32_t* __restrict__ contact_idx, const float* __restrict__ mass, long long n_win, int F,
                   int stride, int last_only, float* __restrict__ out, long long out_ld) {
  const int Fo = last_only ? 1 : F;
  const int CH = 15 * nb;
  const long long n_rows = n_win * Fo;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const long long i = r / Fo;
    const int f = last_only ? F - 1 : (int)(r - i * Fo);
    const float* src = raw + (__ldg(row0 + i) + (long long)f * stride) * raw_ld;
    const float m = __ldg(mass + i);
    for (int ch = lane; ch < CH; ch += 32) {
      // channel → (quantity, body, component)
      int q, w, rem;
      if (ch < 3 * nb) { q = 0; w = 3; rem = ch; }
      else if (ch < 6 * nb) { q = 1; w = 3; rem = ch - 3 * nb; }
      else if (ch < 9 * nb) { q = 2; w = 3; rem = ch - 6 * nb; }
      else { q = 3; w = 6; rem = ch - 9 * nb; }
      const int body = rem / w, comp = rem - body * w;
      const int ci = __ldg(contact_idx + i * nb + body);
      float v = 0.f;
      if (ci >= 0) {
        const int qoff = q == 0 ? 0 : q == 1 ? 3 * nb : q == 2 ? 6 * nb : 9 * nb;
        v = __ldg(src + qoff + ci * w + comp);
        if (q != 0) v = __fdiv_rn(v, m);                    // CoP is not mass-normalised (Dataset.py:251-253)
      }
      out[r * out_ld + ch] = v;
    }
  }
}


// Dict-of-tensors packer (A-4): up to 10 fp32 sources, each [rows, width_k] contiguous, concatenated
// along the channel axis in the given order into one bf16 (and/or fp32) row per (window, frame).
struct PackSrc {
  const float* ptr[10];
  int width[10];
  int offset[11];
  int n;
};
__global__ void __launch_bounds__(kThreads)
pack_inputs_kernel(const PackSrc src, long long n_rows, int F, float* __restrict__ out_f32,
                   __nv_bfloat16* __restrict__ out_bf16, long long fs, long long we, long long col0) {
  const int C = src.offset[src.n];
  const long long n = n_rows * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    int k = 0;
#pragma unroll
    for (int i = 1; i < 10; ++i) k += (i < src.n && c >= src.offset[i]) ? 1 : 0;
    const float v = __ldg(src.ptr[k] + r * src.width[k] + (c - src.offset[k]));
    if (out_f32) out_f32[e] = v;
    if (out_bf16) out_bf16[r * fs + (r / F) * we + col0 + c] = __float2bfloat16_rn(v);
  }
}

static int grid_for(long long items, int per_block) {
  long long need = ceil_div(items, per_block);
  long long cap = (long long)sm_count() * 16;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

}  // namespace ibm

extern "C" int ibm_window_valid_mask(const uint8_t* missing, const int64_t* trial_frame_base, const int32_t* cand_trial,
                                     const int32_t* cand_start, int64_t n_cand, int32_t window_size, int32_t stride,
                                     uint8_t* valid, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(window_size > 0 && stride > 0, "window_valid_mask: window_size and stride must be positive");
  if (n_cand == 0) return IBM_OK;                       // empty dataset → empty index (Dataset.py:134 range(0))
  IBM_CHECK_ARG(missing && trial_frame_base && cand_trial && cand_start && valid && n_cand > 0, "window_valid_mask: null argument");
  window_valid_kernel<<<grid_for(n_cand, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      missing, reinterpret_cast<const long long*>(trial_frame_base), cand_trial, cand_start, n_cand, window_size, stride, valid);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_pack_windows(const float* frames, int64_t frame_ld, int32_t C, const int64_t* win_row0, int64_t n_win,
                                int32_t F, int32_t stride, float* out_f32, void* out_bf16, int64_t bf16_frame_stride,
                                int64_t bf16_win_extra, int64_t bf16_col0, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  if (n_win == 0) return IBM_OK;                        // ragged tail / empty batch
  IBM_CHECK_ARG(frames && win_row0 && n_win > 0 && F > 0 && C > 0 && stride > 0 && frame_ld >= C, "pack_windows: bad argument");
  IBM_CHECK_ARG(out_f32 || out_bf16, "pack_windows: no output requested");
  const bool vec = (frame_ld % 4 == 0) && aligned16(frames) && (!out_f32 || (C % 4 == 0 && aligned16(out_f32))) &&
                   (!out_bf16 || ((bf16_frame_stride % 2 == 0) && (bf16_win_extra % 2 == 0) && (bf16_col0 % 2 == 0) &&
                                  (reinterpret_cast<uintptr_t>(out_bf16) % 4 == 0)));
  const long long n_rows = (long long)n_win * F;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = grid_for(n_rows, kThreads / 32);
  auto* ob = static_cast<__nv_bfloat16*>(out_bf16);
  auto* r0 = reinterpret_cast<const long long*>(win_row0);
  if (vec) pack_windows_kernel<true><<<grid, kThreads, 0, s>>>(frames, frame_ld, C, r0, n_rows, F, stride, out_f32, ob,
                                                              bf16_frame_stride, bf16_win_extra, bf16_col0);
  else pack_windows_kernel<false><<<grid, kThreads, 0, s>>>(frames, frame_ld, C, r0, n_rows, F, stride, out_f32, ob,
                                                            bf16_frame_stride, bf16_win_extra, bf16_col0);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_pack_labels(const float* raw, int64_t raw_ld, int32_t nb, const int64_t* win_row0,
                               const int32_t* contact_idx, const float* mass, int64_t n_win, int32_t F, int32_t stride,
                               int32_t last_frame_only, float* out_rows, int64_t out_ld, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  if (n_win == 0) return IBM_OK;
  IBM_CHECK_ARG(raw && win_row0 && contact_idx && mass && out_rows && n_win > 0 && F > 0 && nb > 0 && stride > 0,
                "pack_labels: bad argument");
  IBM_CHECK_ARG(raw_ld >= 15 * nb && out_ld >= 15 * nb, "pack_labels: leading dimensions too small");
  const long long n = (long long)n_win * (last_frame_only ? 1 : F) * 32;      // one warp per output row
  pack_labels_kernel<<<grid_for(n, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      raw, raw_ld, nb, reinterpret_cast<const long long*>(win_row0), contact_idx, mass, n_win, F, stride, last_frame_only,
      out_rows, out_ld);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

// ---- channel-major windows -> frame rows (TransformerBaseline.forward, TransformerBaseline.py:108-126) ----------------
// The transformer's inputs arrive as (B, C_i, T) tensors (channel-major, concatenated on dim 1 and transposed by the
// reference); the layer stack wants bf16 rows [B*T, ld] with the learned temporal embedding of frame t appended.  One
// block = 64 frames of one window: the (C, 64) slab goes through shared memory so that both the reads (along T) and
// the writes (along the row) are coalesced.
namespace ibm {
constexpr int kCmFrames = 64;
struct ChanSrc {
  const float* ptr[8];
  int32_t ch[8];
  int32_t first[9];      // first output column of each source
  int32_t n;
};
__global__ void __launch_bounds__(kThreads)
pack_channel_major_kernel(ChanSrc src, int T, const float* __restrict__ emb, int E, __nv_bfloat16* __restrict__ out,
                          long long ld) {
  extern __shared__ float tile[];                       // [C][kCmFrames + 1]
  const long long b = blockIdx.x;
  const int t0 = blockIdx.y * kCmFrames;
  const int nt = min(kCmFrames, T - t0);
  const int C = src.first[src.n];
  for (int s = 0; s < src.n; ++s) {
    const float* p = src.ptr[s] + b * (long long)src.ch[s] * T + t0;
    for (int i = threadIdx.x; i < src.ch[s] * kCmFrames; i += kThreads) {
      const int c = i / kCmFrames, f = i - c * kCmFrames;
      if (f < nt) tile[(src.first[s] + c) * (kCmFrames + 1) + f] = __ldg(p + (long long)c * T + f);
    }
  }
  __syncthreads();
  const int pairs = (int)(ld >> 1);
  for (int i = threadIdx.x; i < nt * pairs; i += kThreads) {
    const int f = i / pairs, c = 2 * (i - f * pairs);
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int cc = c + j;
      v[j] = cc < C ? tile[cc * (kCmFrames + 1) + f] : (cc < C + E ? __ldg(emb + (long long)(t0 + f) * E + (cc - C)) : 0.f);
    }
    *reinterpret_cast<uint32_t*>(out + (b * T + t0 + f) * ld + c) = pack_bf16x2(v[0], v[1]);
  }
}
// Pre-packed analysis stream (BASELINE configs[4]): rows that were converted ONCE on the host to frame-major bf16
// ([B*T, ld_src], the C kinematic channels of a frame contiguous) are expanded into the transformer's input rows
// [B*T, ld] = [C channels | E temporal-embedding columns of frame t | zero pad] — bit-identical to what
// pack_channel_major_kernel writes from the fp32 channel-major tensors (same RNE bf16 values) — and, optionally, columns
// [vcol0, vcol0 + 3) go to an 8-wide value buffer (the CoM accelerations the SimpleAttention blends, TransformerBaseline.py:135).
// One thread per 2 output columns; reads and writes are both row-contiguous.
__global__ void __launch_bounds__(kThreads)
expand_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long ld_src, int C, int T, const float* __restrict__ emb, int E,
                        __nv_bfloat16* __restrict__ out, long long ld, __nv_bfloat16* __restrict__ vout, int vcol0, long long n_rows) {
  const int pairs = (int)(ld >> 1);
  const long long total = n_rows * pairs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / pairs;
    const int c = 2 * (int)(i - r * pairs);
    const int t = (int)(r % T);
    uint32_t w;
    if (c + 1 < C) {
      w = *reinterpret_cast<const uint32_t*>(src + r * ld_src + c);
    } else {
      float v[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int cc = c + j;
        v[j] = cc < C ? __bfloat162float(src[r * ld_src + cc]) : (cc < C + E ? __ldg(emb + (long long)t * E + (cc - C)) : 0.f);
      }
      w = pack_bf16x2(v[0], v[1]);
    }
    *reinterpret_cast<uint32_t*>(out + r * ld + c) = w;
    if (vout != nullptr && c < 8) {                       // threads of output columns 0..7 also write the 8-wide value row
      float v[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) v[j] = (c + j < 3) ? __bfloat162float(src[r * ld_src + vcol0 + c + j]) : 0.f;
      *reinterpret_cast<uint32_t*>(vout + r * 8 + c) = pack_bf16x2(v[0], v[1]);
    }
  }
}
}  // namespace ibm

extern "C" int ibm_pack_channel_major(const void* const* h_src, const int32_t* h_channels, int32_t n_src, int64_t B, int32_t T,
                                      const float* emb, int32_t E, void* out_bf16, int64_t ld, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(h_src && h_channels && n_src > 0 && n_src <= 8 && B > 0 && T > 0 && out_bf16, "pack_channel_major: bad argument");
  IBM_CHECK_ARG(E >= 0 && (E == 0 || emb), "pack_channel_major: embedding width without a table");
  ChanSrc src;
  src.n = n_src;
  src.first[0] = 0;
  for (int i = 0; i < 8; ++i) {
    src.ptr[i] = i < n_src ? static_cast<const float*>(h_src[i]) : nullptr;
    src.ch[i] = i < n_src ? h_channels[i] : 0;
    IBM_CHECK_ARG(i >= n_src || (src.ptr[i] && src.ch[i] > 0), "pack_channel_major: null source %d", i);
    src.first[i + 1] = src.first[i] + src.ch[i];
  }
  const int C = src.first[n_src];
  IBM_CHECK_ARG(ld % 8 == 0 && ld >= C + E && C <= 256, "pack_channel_major: ld must be a multiple of 8 and >= C + E (C <= 256)");
  IBM_CHECK_ARG(B < (1ll << 31), "pack_channel_major: too many windows for one launch");
  const size_t smem = (size_t)C * (kCmFrames + 1) * sizeof(float);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    IBM_CHECK_CUDA(cudaFuncSetAttribute(pack_channel_major_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((unsigned)B, (unsigned)ceil_div(T, kCmFrames));
  pack_channel_major_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(src, T, emb, E,
                                                                                         static_cast<__nv_bfloat16*>(out_bf16), ld);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_pack_inputs(const void* const* h_src, const int32_t* h_widths, int32_t n_src, int64_t n_rows, int32_t F,
                               float* out_f32, void* out_bf16, int64_t bf16_frame_stride, int64_t bf16_win_extra,
                               int64_t bf16_col0, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  if (n_rows == 0) return IBM_OK;
  IBM_CHECK_ARG(h_src && h_widths && n_src > 0 && n_src <= 10 && n_rows > 0 && F > 0, "pack_inputs: bad argument");
  IBM_CHECK_ARG(out_f32 || out_bf16, "pack_inputs: no output requested");
  PackSrc src;
  src.n = n_src;
  src.offset[0] = 0;
  for (int i = 0; i < 10; ++i) {
    src.ptr[i] = i < n_src ? static_cast<const float*>(h_src[i]) : nullptr;
    src.width[i] = i < n_src ? h_widths[i] : 0;
    IBM_CHECK_ARG(i >= n_src || (src.ptr[i] && src.width[i] > 0), "pack_inputs: null source %d", i);
    src.offset[i + 1] = src.offset[i] + src.width[i];
  }
  pack_inputs_kernel<<<grid_for(n_rows * src.offset[n_src], kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      src, n_rows, F, out_f32, static_cast<__nv_bfloat16*>(out_bf16), bf16_frame_stride, bf16_win_extra, bf16_col0);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}

extern "C" int ibm_expand_rows_bf16(const void* src_bf16, int64_t ld_src, int32_t C, int64_t n_rows, int32_t T, const float* emb,
                                    int32_t E, void* out_bf16, int64_t ld, void* v_out_bf16, int32_t v_col0, void* stream) {
  using namespace ibm;
  IBM_CHECK_ARCH();
  IBM_CHECK_ARG(src_bf16 && out_bf16 && n_rows > 0 && T > 0 && C > 0 && n_rows % T == 0, "expand_rows_bf16: bad argument");
  IBM_CHECK_ARG(E >= 0 && (E == 0 || emb), "expand_rows_bf16: embedding width without a table");
  IBM_CHECK_ARG(ld % 2 == 0 && ld >= C + E && ld_src % 2 == 0 && ld_src >= C, "expand_rows_bf16: ld must be even and >= C + E, ld_src even and >= C");
  IBM_CHECK_ARG(v_out_bf16 == nullptr || (v_col0 >= 0 && v_col0 + 3 <= C), "expand_rows_bf16: value columns outside the row");
  expand_rows_bf16_kernel<<<grid_for(n_rows * (ld / 2), kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src_bf16), ld_src, C, T, emb, E, static_cast<__nv_bfloat16*>(out_bf16), ld,
      static_cast<__nv_bfloat16*>(v_out_bf16), v_col0, n_rows);
  IBM_LAUNCH_CHECK();
  return IBM_OK;
}
