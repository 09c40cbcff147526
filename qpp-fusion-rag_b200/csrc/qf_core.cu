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
  static bool ok[64] = {};                  // devices already verified (the answer cannot change)
  if (dev >= 0 && dev < 64 && ok[dev]) return QF_OK;
  const int rc = qf_device_check(dev);
  if (rc == QF_OK && dev >= 0 && dev < 64) ok[dev] = true;
  return rc;
}

// Optional device-side timing (qf_profile_*): up to PROF_MAX event pairs per phase since the last enable; the read
// returns their average.
constexpr int PROF_MAX = 512;
struct ProfPhase {
  cudaEvent_t ev[PROF_MAX][2];
  int created = 0, n = 0;
  bool open = false;
};
static bool g_prof_on = false;
static ProfPhase g_prof[QF_PROF_PHASES];

void prof_mark(int phase, int which, cudaStream_t st) {
  if (!g_prof_on || phase < 0 || phase >= QF_PROF_PHASES) return;
  ProfPhase& ph = g_prof[phase];
  if (ph.n >= PROF_MAX) return;
  if (which == 0) {
    if (ph.n >= ph.created) {
      if (cudaEventCreate(&ph.ev[ph.n][0]) != cudaSuccess || cudaEventCreate(&ph.ev[ph.n][1]) != cudaSuccess) return;
      ph.created = ph.n + 1;
    }
    cudaEventRecord(ph.ev[ph.n][0], st);
    ph.open = true;
  } else if (ph.open) {
    cudaEventRecord(ph.ev[ph.n][1], st);
    ph.open = false;
    ++ph.n;
  }
}

int device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return n;
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) cached[dev] = n;
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
  if (on)
    for (int i = 0; i < QF_PROF_PHASES; ++i) {
      qf::g_prof[i].n = 0;
      qf::g_prof[i].open = false;
    }
  return QF_OK;
}

extern "C" int qf_profile_read(int phase, float* ms) {
  QF_REQUIRE(ms != nullptr && phase >= 0 && phase < QF_PROF_PHASES, QF_E_INVALID, "qf_profile_read: bad argument");
  qf::ProfPhase& ph = qf::g_prof[phase];
  QF_REQUIRE(ph.n > 0, QF_E_INVALID, "qf_profile_read: phase %d was not recorded", phase);
  QF_CUDA_TRY(cudaEventSynchronize(ph.ev[ph.n - 1][1]));
  double sum = 0.0;
  for (int i = 0; i < ph.n; ++i) {
    float t = 0.f;
    QF_CUDA_TRY(cudaEventElapsedTime(&t, ph.ev[i][0], ph.ev[i][1]));
    sum += t;
  }
  *ms = (float)(sum / ph.n);
  return QF_OK;
}
