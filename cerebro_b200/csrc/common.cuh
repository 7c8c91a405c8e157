// Shared helpers for the C-ABI translation units (error plumbing, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/cerebro_b200.h"

namespace cb {

char* tls_error_buffer();  // defined in capi.cu
inline int fail(int code, const char* fmt, ...) {
  char* buf = tls_error_buffer();
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return cb::fail(CB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                      __FILE__, __LINE__);                                                    \
  } while (0)

#define CB_LAUNCH_CHECK()                                                                     \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess)                                                                    \
      return cb::fail(CB_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),   \
                      __FILE__, __LINE__);                                                    \
  } while (0)

// Select `device`, verify it is a Blackwell (sm_100) part; fills sm count.
int select_device(int device, int* sm_count);

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

constexpr unsigned FULL = 0xffffffffu;

// Host-side wait for a stream, used by every blocking (host-pointer) entry point.  cudaStreamSynchronize spins, which is the
// lowest latency when the calling thread has a core of its own (3-query rule 0.64 ms against 1.02 ms) -- and wastes the cores
// when it has not: the bench at 8 ranks on a 32-core box runs six host threads per rank (descriptor x3, search, verifier,
// main) on 4 cores; measured there, sleeping gave 89 k keyframes/s e2e against 86 k spinning.  Policy (CB_SYNC=spin|block
// overrides): sleep on an interrupt-driven event when the machine has fewer than 8 online cores per visible GPU, spin otherwise.
cudaError_t sync_stream(cudaStream_t st);  // defined in capi.cu

}  // namespace cb
