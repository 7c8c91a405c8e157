// Library-wide pieces of the C ABI: version, error string, device selection.
#include "common.cuh"

#include <stdlib.h>
#include <unistd.h>

namespace cb {

namespace {
bool decide_blocking_sync() {
  if (const char* e = getenv("CB_SYNC")) {
    if (!strcmp(e, "block")) return true;
    if (!strcmp(e, "spin")) return false;
  }
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess) {
    cudaGetLastError();
    n_dev = 1;
  }
  const long cores = sysconf(_SC_NPROCESSORS_ONLN);
  return cores > 0 && cores < 8L * (n_dev > 0 ? n_dev : 1);
}
}  // namespace

cudaError_t sync_stream(cudaStream_t st) {
  static const bool blocking = decide_blocking_sync();
  if (!blocking) return cudaStreamSynchronize(st);
  // one interrupt-driven event per (thread, device); the waiting thread sleeps instead of spinning on a core it shares
  constexpr int kMaxDev = 64;
  static thread_local cudaEvent_t ev[kMaxDev] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDev) return cudaStreamSynchronize(st);
  if (!ev[dev]) {
    e = cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  e = cudaEventRecord(ev[dev], st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev[dev]);
}

char* tls_error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int select_device(int device, int* sm_count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(CB_ENODEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(CB_EINVAL, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(CB_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(CB_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  if (sm_count) *sm_count = prop.multiProcessorCount;
  return CB_OK;
}

}  // namespace cb

extern "C" {

int cb_version(void) { return 2; }

const char* cb_last_error(void) { return cb::tls_error_buffer(); }

int cb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void cb_ransac_params_default(cb_ransac_params* p) {
  if (!p) return;
  p->error_thresh = 0.03;      // DlsPnpWithRansac.cpp:208
  p->min_inlier_ratio = 0.7;   // :209
  p->max_iterations = 50;      // :210
  p->min_iterations = 5;       // :211
  p->use_mle = 1;              // :212
  p->failure_probability = 0.01;
  p->adaptive = 1;
  p->seed = 0;
}

}  // extern "C"
