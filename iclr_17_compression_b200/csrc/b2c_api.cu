// libb2c: version, error reporting, device check, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "b2c_common.cuh"

namespace b2c {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return B2C_ECUDA;
  }
  return B2C_OK;
}

}  // namespace b2c

extern "C" {

int b2c_version(void) { return 100; }

const char* b2c_last_error(void) { return b2c::g_err; }

uint64_t b2c_launch_count(void) { return b2c::g_launches.load(std::memory_order_relaxed); }

int b2c_check_device(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    b2c::set_error("cudaGetDevice failed: %s (no CUDA device; libb2c has no CPU fallback)", cudaGetErrorString(e));
    return B2C_ECUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    b2c::set_error("device %d is sm_%d%d; libb2c is built for sm_100a only (no fallback path)", dev, major, minor);
    return B2C_EARCH;
  }
  return B2C_OK;
}

}  // extern "C"
