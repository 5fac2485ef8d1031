// Error reporting, version and device queries behind include/mri_b200.h.
#include <stdarg.h>

#include "common.cuh"

namespace mri {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MRI_OK;
  return fail(MRI_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace mri

extern "C" int mri_version(void) { return MRI_B200_VERSION; }
extern "C" const char* mri_last_error(void) { return mri::error_buffer(); }
extern "C" int mri_sm_count(void) { return mri::sm_count(); }
