// qf_dense_rescore_f32: float32 re-scoring of an over-fetched candidate list.
//
// The reference searches a float32 faiss.IndexFlatIP (src/retrievers/bge.py:188-194); qf_dense_topk scores bf16-rounded
// operands on the tensor cores (|score error| <= 4e-3 for unit-norm 768-d rows, measured 2.5e-4), so membership and
// order can differ from the reference near the rank-k boundary.  Re-scoring the top (k + margin) candidates against the
// float32 corpus rows removes that: out[q][j] = <queries[q], corpus[cand_ids[q][j] - id_offset]> with float32 FMAs.
// One warp per (query, candidate) pair: 16-byte loads of the corpus row (each row is read exactly once per candidate:
// HBM-bound, (k + margin) * D * 4 bytes per query), per-lane partial sums in a fixed order, fixed shuffle tree.
#include "qf_common.cuh"

namespace {

constexpr int RS_WARPS = 8;

__global__ void __launch_bounds__(RS_WARPS * 32) rescore_kernel(const float* __restrict__ queries,
                                                                 const float* __restrict__ corpus, int Q, long long N, int D,
                                                                 const int32_t* __restrict__ cand, int kc,
                                                                 long long id_offset, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long pair = (long long)blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
  if (pair >= (long long)Q * kc) return;
  const int q = (int)(pair / kc);
  const long long id = (long long)cand[pair] - id_offset;
  if (id < 0 || id >= N) {                       // padding (-1) or an id outside this shard
    if (lane == 0) out[pair] = -INFINITY;
    return;
  }
  const float4* qr = reinterpret_cast<const float4*>(queries + (size_t)q * D);
  const float4* cr = reinterpret_cast<const float4*>(corpus + (size_t)id * D);
  float acc = 0.0f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 a = __ldg(qr + i), b = __ldg(cr + i);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[pair] = acc;
}

}  // namespace

extern "C" int qf_dense_rescore_f32(const float* queries, const float* corpus, int Q, long long N, int D,
                                    const int32_t* cand_ids, int kc, long long id_offset, float* out_scores,
                                    void* stream) {
  QF_REQUIRE(Q >= 0 && N >= 1 && D >= 4 && kc >= 1, QF_E_INVALID, "qf_dense_rescore_f32: bad sizes Q=%d N=%lld D=%d kc=%d",
             Q, N, D, kc);
  QF_REQUIRE(D % 4 == 0, QF_E_UNSUPPORTED, "qf_dense_rescore_f32: D=%d must be a multiple of 4", D);
  if (Q == 0) return QF_OK;
  QF_REQUIRE(queries && corpus && cand_ids && out_scores, QF_E_INVALID, "qf_dense_rescore_f32: null pointer argument");
  QF_REQUIRE(((reinterpret_cast<uintptr_t>(queries) | reinterpret_cast<uintptr_t>(corpus)) & 15) == 0, QF_E_INVALID,
             "qf_dense_rescore_f32: queries/corpus must be 16-byte aligned");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const long long pairs = (long long)Q * kc;
  const long long blocks = (pairs + RS_WARPS - 1) / RS_WARPS;
  QF_REQUIRE(blocks <= 0x7fffffffLL, QF_E_UNSUPPORTED, "qf_dense_rescore_f32: too many candidate pairs");
  rescore_kernel<<<(unsigned)blocks, RS_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      queries, corpus, Q, N, D, cand_ids, kc, id_offset, out_scores);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}
