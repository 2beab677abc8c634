// Library plumbing: error buffer, architecture gate, workspace size.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ibm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_arch_ok[64];   // 0 unknown, 1 ok, 2 bad
static int g_sms[64];

int check_arch() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device — libibm_b200 has no CPU fallback)", cudaGetErrorString(e));
    return IBM_E_CUDA;
  }
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_arch_ok[dev] == 0) {
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_sms[dev] = sms;
    g_arch_ok[dev] = (major == 10) ? 1 : 2;
  }
  if (g_arch_ok[dev] != 1) {
    set_error("device %d is not compute capability 10.x; libibm_b200 targets sm_100a only (no fallback)", dev);
    return IBM_E_ARCH;
  }
  return IBM_OK;
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (g_sms[dev] == 0) cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
  return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

// ---- walk order of the big row-streaming kernels (ibm_set_walk_order) --------------------------------------------
// mode 0: every kernel walks its rows / tiles in ascending order.  mode 1 ("zigzag"): big launches alternate between
// ascending and descending, so a kernel starts on the rows its producer wrote LAST — the part of the producer's output that
// is still in the 126 MB L2 — instead of on rows that were evicted hundreds of megabytes ago.  mode 2: every launch
// alternates, whatever its size (parity tests of the descending paths on small shapes).
static int g_walk_mode = -1;
static unsigned g_walk_count = 0;

int next_walk_reverse(int64_t bytes_streamed) {
  if (g_walk_mode < 0) {
    const char* e = getenv("IBM_WALK_ORDER");
    g_walk_mode = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  if (g_walk_mode == 0 || (g_walk_mode == 1 && bytes_streamed < (int64_t)(48ll << 20))) return 0;
  return (int)(g_walk_count++ & 1u);
}

}  // namespace ibm

extern "C" {

int ibm_version(void) { return 100; }

int ibm_set_walk_order(int32_t mode) {
  if (mode < 0 || mode > 2) {
    ibm::set_error("ibm_set_walk_order: mode must be 0 (ascending), 1 (big launches alternate) or 2 (all launches alternate), got %d", mode);
    return IBM_E_ARG;
  }
  ibm::g_walk_mode = mode;
  ibm::g_walk_count = 0;
  return IBM_OK;
}

size_t ibm_last_error(char* buf, size_t cap) {
  size_t n = strlen(ibm::g_err);
  if (buf && cap) {
    size_t m = n < cap - 1 ? n : cap - 1;
    memcpy(buf, ibm::g_err, m);
    buf[m] = 0;
  }
  return n;
}

int ibm_device_check(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) {
    ibm::set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
    return IBM_E_CUDA;
  }
  if (major != 10) {
    ibm::set_error("device %d has compute capability major %d, need 10 (sm_100a)", device, major);
    return IBM_E_ARCH;
  }
  return IBM_OK;
}

// 1024 partial blocks x 40 floats + counter, rounded up
size_t ibm_workspace_bytes(void) { return 256 * 1024; }

}  // extern "C"
