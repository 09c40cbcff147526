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

static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[QF_PROF_PHASES][2];
static bool g_prof_init = false, g_prof_rec[QF_PROF_PHASES];

void prof_mark(int phase, int which, cudaStream_t st) {
  if (!g_prof_on || phase < 0 || phase >= QF_PROF_PHASES) return;
  if (!g_prof_init) {
    for (int i = 0; i < QF_PROF_PHASES; ++i) {
      cudaEventCreate(&g_prof_ev[i][0]);
      cudaEventCreate(&g_prof_ev[i][1]);
      g_prof_rec[i] = false;
    }
    g_prof_init = true;
  }
  cudaEventRecord(g_prof_ev[phase][which], st);
  if (which == 1) g_prof_rec[phase] = true;
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

extern "C" int qf_profile_enable(int on) {
  qf::g_prof_on = on != 0;
  return QF_OK;
}

extern "C" int qf_profile_read(int phase, float* ms) {
  QF_REQUIRE(ms != nullptr && phase >= 0 && phase < QF_PROF_PHASES, QF_E_INVALID, "qf_profile_read: bad argument");
  QF_REQUIRE(qf::g_prof_init && qf::g_prof_rec[phase], QF_E_INVALID, "qf_profile_read: phase %d was not recorded", phase);
  QF_CUDA_TRY(cudaEventSynchronize(qf::g_prof_ev[phase][1]));
  QF_CUDA_TRY(cudaEventElapsedTime(ms, qf::g_prof_ev[phase][0], qf::g_prof_ev[phase][1]));
  return QF_OK;
}
