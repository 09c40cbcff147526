// Types shared by the two rank-fusion kernels (qf_fuse_stream.cu: streaming hash-accumulate + bucket ranking, the
// common case; qf_fuse.cu: fully staged general kernel that also finishes the queries the streaming kernel defers).
#pragma once
#include "qf_common.cuh"

namespace qf {

constexpr int FUSE_MAX_R = 32;
constexpr uint32_t FUSE_FLAG_DEBUG_FULLSORT = 0x100u;   // test hook: general kernel, force the 64-bit sort path
constexpr uint32_t FUSE_FLAG_DEBUG_GENERAL = 0x200u;    // test hook: skip the streaming kernels
constexpr uint32_t FUSE_FLAG_GATHER = 0x400u;           // opt-in: the deposit + gather kernel (qf_fuse_gather.cu) instead of the streaming one
constexpr int FUSE_DEFERRED = -1;                       // out_lens marker: query left to the general kernel

struct FuseParams {
  const int32_t* doc_ids;
  const void* scores;
  const int32_t* ranks;
  const int32_t* lens;
  const double* weights;
  int Q, R, K;
  int method, norm;
  double rrf_k;
  uint32_t flags;
  int32_t* out_ids;
  double* out_scores;
  unsigned short* out_counts;
  unsigned short* out_pos;
  int32_t* out_lens;
  int32_t* out_status;
  int fixup;          // general kernel only: process just the queries whose out_lens == FUSE_DEFERRED
};

template <int ST> struct ScoreT;
template <> struct ScoreT<QF_SCORE_F64>   { using type = double; };
template <> struct ScoreT<QF_SCORE_F32>   { using type = float; };
template <> struct ScoreT<QF_SCORE_MICRO> { using type = int; };

__device__ __forceinline__ double raw_to_f64(double v) { return v; }
__device__ __forceinline__ double raw_to_f64(float v) { return (double)v; }
__device__ __forceinline__ double raw_to_f64(int v) { return micro_to_f64(v); }

// Streaming kernel (qf_fuse_stream.cu).  fuse_stream_supported: does the request fit its envelope at all;
// fuse_stream_launch: enqueue it.  Queries it cannot finish (union too large, duplicate doc id inside a list,
// long runs of exactly tied scores, hash-probe overflow) get out_lens[q] = FUSE_DEFERRED.
bool fuse_stream_supported(const FuseParams& p);
int fuse_stream_launch(const FuseParams& p, int score_type, cudaStream_t stream);

// Gather kernel (qf_fuse_gather.cu, round 2): same contract as the streaming kernel, smaller envelope (R <= 8,
// K <= 1024, 4-byte scores).  Measured on B200 (profiles/r2c_fuse_gather_ncu_summary.txt): 20 % fewer shared-memory
// wavefronts but 12 % more instructions than the streaming kernel and 10 % SLOWER -- the fusion kernels are bound by
// instruction issue, not by the LSU pipe -- so it is opt-in (FUSE_FLAG_GATHER / QF_FUSE_GATHER=1), kept with full
// parity coverage as the measured alternative.
bool fuse_gather_supported(const FuseParams& p, int score_type);
int fuse_gather_launch(const FuseParams& p, int score_type, cudaStream_t stream);

}  // namespace qf
