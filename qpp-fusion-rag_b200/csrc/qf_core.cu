// Error plumbing, device checks and ABI version of libqppfusion.
#include <stdarg.h>

#include "qf_common.cuh"

namespace qf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device: %s (libqppfusion has no CPU fallback)", cudaGetErrorString(e));
    return QF_E_NO_DEVICE;
  }
  return qf_device_check(dev);
}

int sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

}  // namespace qf

extern "C" const char* qf_last_error(void) { return qf::g_err; }

extern "C" int qf_abi_version(void) { return QF_ABI_VERSION; }

extern "C" int qf_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    qf::set_error("no CUDA device: %s (libqppfusion has no CPU fallback)",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return QF_E_NO_DEVICE;
  }
  QF_REQUIRE(device >= 0 && device < n, QF_E_INVALID, "device %d out of range [0,%d)", device, n);
  int major = 0, minor = 0;
  QF_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  QF_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    qf::set_error("device %d is sm_%d%d; libqppfusion is built for sm_100a only and has no fallback", device,
                  major, minor);
    return QF_E_NO_DEVICE;
  }
  return QF_OK;
}
