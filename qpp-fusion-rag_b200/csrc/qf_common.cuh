// Shared device/host helpers for libqppfusion (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "qppfusion.h"

namespace qf {

// ---- host side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_device();                       // QF_OK or QF_E_NO_DEVICE (message set)
int  sm_count();
int  device_index();                    // current CUDA device, clamped to [0, 63] (index of per-device one-time state)
void prof_mark(int phase, int which, cudaStream_t st);   // CUDA-event pair around a kernel (qf_profile_*)

#define QF_CUDA_TRY(expr)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      qf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return QF_E_CUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define QF_REQUIRE(cond, code, ...)     \
  do {                                  \
    if (!(cond)) {                      \
      qf::set_error(__VA_ARGS__);       \
      return (code);                    \
    }                                   \
  } while (0)

// ---- device helpers -----------------------------------------------------------------------------
// Order-preserving map double -> uint64 (ascending).  -0.0 must be canonicalised by the caller.
__device__ __forceinline__ unsigned long long ord_f64(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// The double that Python's float("d.dddddd") parses to, from the integer number of micro-units.
// i / 1e6 is an IEEE division of two exactly representable values, hence the correctly rounded
// double nearest to the decimal -- exactly what strtod returns.
// Division-free (Markstein correction): q0 = a*1e-6; r = fma(-1e6, q0, a); q = fma(r, 1e-6, q0).  Verified
// EXHAUSTIVELY equal to a/1e6 for every int32 a (oracle/qf_oracle.c qfo_check_micro_division,
// tests/test_oracle_c.py).
__device__ __forceinline__ double micro_to_f64(int v) {
  const double a = (double)v;
  const double q0 = __dmul_rn(a, 1.0e-6);
  const double r = __fma_rn(-1.0e6, q0, a);
  return __fma_rn(r, 1.0e-6, q0);
}

template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace qf
