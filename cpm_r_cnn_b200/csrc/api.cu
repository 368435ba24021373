// Library-level entry points of libcpm_ops.so: error reporting, launch accounting, device binding.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace cpm {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_device_ptr(const void* p, const char* what) {
  if (p == nullptr) {
    set_error("%s is NULL", what);
    return CPM_ERR_INVALID_ARG;
  }
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cudaPointerGetAttributes failed: %s", what, cudaGetErrorString(e));
    return CPM_ERR_CUDA;
  }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) {
    // the reference's "input must be a CUDA tensor" (ROIAlign_cuda.cu:376); there is no CPU path here
    set_error("%s must be a CUDA device pointer (no CPU fallback in cpm_ops)", what);
    return CPM_ERR_INVALID_ARG;
  }
  return CPM_OK;
}

}  // namespace cpm

extern "C" {

const char* cpm_last_error(void) { return cpm::g_error; }

int cpm_version(void) { return 100; }

uint64_t cpm_launch_count(void) { return cpm::g_launches.load(std::memory_order_relaxed); }

int cpm_set_device(int device) {
  CPM_CHECK_CUDA(cudaSetDevice(device));
  return CPM_OK;
}

}  // extern "C"
