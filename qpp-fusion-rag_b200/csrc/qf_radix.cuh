// Block-wide stable LSD radix pass in shared memory (8-bit digits), shared by qf_fuse and qf_dense.
#pragma once
#include "qf_common.cuh"

namespace qf {

using u16 = unsigned short;
using u32 = unsigned int;
using u64 = unsigned long long;

// One stable LSD radix pass over `n_pad` (multiple of 32) elements.
// digit(e) is supplied by `dig`; payload = (optional key32, idx16).  Ranking inside a warp uses
// __match_any_sync; per-warp digit counters live in shared memory as whist[digit][warp] (only the lowest
// lane of each digit group updates a counter, so no atomics); the digit-major scan that turns them into
// scatter bases reads a digit's NW counters as one vector.
// Returns false (and moves nothing) when every element has the same digit.
template <int NT, int MAXROUNDS, bool CARRY_KEY, typename DigitFn>
__device__ __forceinline__ bool radix_pass(DigitFn dig, int n_pad, const u32* keyIn, u32* keyOut, const u16* idxIn,
                                           u16* idxOut, u16* whist, int* warp_tot, int* single_flag) {
  constexpr int NW = NT / 32;
  static_assert(NW % 2 == 0 && NW <= 32, "NW must be even and <= 32");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rounds_total = n_pad >> 5;
  const int rounds_per_warp = (rounds_total + NW - 1) / NW;
  const int r0 = warp * rounds_per_warp;
  const int r1 = min(rounds_total, r0 + rounds_per_warp);

  {
    u32* w32 = reinterpret_cast<u32*>(whist);
    for (int i = tid; i < NW * 128; i += NT) w32[i] = 0u;
  }
  if (tid == 0) *single_flag = 0;
  __syncthreads();

  u16 lrank[MAXROUNDS];
  u16 ldig[MAXROUNDS];
#pragma unroll
  for (int k = 0; k < MAXROUNDS; ++k) {
    const int rd = r0 + k;
    if (rd < r1) {
      const int e = (rd << 5) + lane;
      const u32 d = dig(e);
      const u32 peers = __match_any_sync(0xffffffffu, d);
      const u32 lt = peers & ((1u << lane) - 1u);
      u32 base = 0;
      if (lt == 0) {
        base = whist[d * NW + warp];
        whist[d * NW + warp] = (u16)(base + __popc(peers));
      }
      __syncwarp();
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      lrank[k] = (u16)(base + __popc(lt));
      ldig[k] = (u16)d;
    }
  }
  __syncthreads();

  // digit-major exclusive scan: thread t < SCAN_T owns digits [t*DPT, (t+1)*DPT)
  constexpr int DPT = NT >= 256 ? 1 : 256 / NT;
  constexpr int SCAN_T = 256 / DPT;
  {
    u32 packed[DPT][NW / 2];
    int dtot[DPT];
    int tot = 0;
    if (tid < SCAN_T) {
#pragma unroll
      for (int j = 0; j < DPT; ++j) {
        const int d = tid * DPT + j;
        const u32* row = reinterpret_cast<const u32*>(whist + d * NW);
        int run = 0;
#pragma unroll
        for (int w2 = 0; w2 < NW / 2; ++w2) {
          const u32 v = row[w2];
          const int c0 = (int)(v & 0xFFFFu), c1 = (int)(v >> 16);
          packed[j][w2] = (u32)run | ((u32)(run + c0) << 16);
          run += c0 + c1;
        }
        dtot[j] = run;
        tot += run;
        if (run == n_pad) *single_flag = 1;   // every element has this digit: the pass is the identity
      }
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (tid < SCAN_T && lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (tid < SCAN_T) {
      int excl = incl - tot;
      for (int w = 0; w < warp; ++w) excl += warp_tot[w];
#pragma unroll
      for (int j = 0; j < DPT; ++j) {
        const int d = tid * DPT + j;
        u32* row = reinterpret_cast<u32*>(whist + d * NW);
        const u32 add = (u32)excl * 0x00010001u;
#pragma unroll
        for (int w2 = 0; w2 < NW / 2; ++w2) row[w2] = packed[j][w2] + add;   // no carry: totals < 65536
        excl += dtot[j];
      }
    }
    __syncthreads();
  }
  if (*single_flag) {
    __syncthreads();
    return false;   // nothing moves; caller keeps the current buffers
  }
#pragma unroll
  for (int k = 0; k < MAXROUNDS; ++k) {
    const int rd = r0 + k;
    if (rd < r1) {
      const int e = (rd << 5) + lane;
      const int pos = whist[(int)ldig[k] * NW + warp] + lrank[k];
      if (CARRY_KEY) keyOut[pos] = keyIn[e];
      idxOut[pos] = idxIn[e];
    }
  }
  __syncthreads();
  return true;
}

}  // namespace qf
