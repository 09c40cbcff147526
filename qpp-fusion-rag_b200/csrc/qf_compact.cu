// qf_compact_fused: padded [Q, stride] fusion outputs -> CSR (offsets[Q+1], ids[total], scores[total]).
// The fused run of a query is out_lens[q] <= R*K entries; compacting before the device->host copy (or before
// writing a .res file) moves |union| instead of R*K entries per query.
#include "qf_common.cuh"

namespace {

constexpr int SCAN_NT = 256;

// block b sums lens[b*256 .. +256)
__global__ void __launch_bounds__(SCAN_NT) block_sums(const int32_t* lens, int Q, long long* bsum) {
  __shared__ long long sm[SCAN_NT / 32];
  const int i = blockIdx.x * SCAN_NT + threadIdx.x;
  long long v = i < Q ? (long long)lens[i] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < SCAN_NT / 32; ++w) t += sm[w];
    bsum[blockIdx.x] = t;
  }
}

// single block: exclusive scan of the block sums in place; total -> offsets[Q]
__global__ void __launch_bounds__(SCAN_NT) scan_block_sums(long long* bsum, int nb, long long* offsets, int Q) {
  __shared__ long long carry;
  __shared__ long long wsum[SCAN_NT / 32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_NT) {
    const int i = base + threadIdx.x;
    const long long v = i < nb ? bsum[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    long long wb = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wb += wsum[w];
    const long long c = carry;
    if (i < nb) bsum[i] = c + wb + incl - v;
    __syncthreads();
    if (threadIdx.x == SCAN_NT - 1) carry = c + wb + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[Q] = carry;
}

// offsets[q] = block base + exclusive scan inside the block
__global__ void __launch_bounds__(SCAN_NT) finish_offsets(const int32_t* lens, int Q, const long long* bsum, long long* offsets) {
  __shared__ long long wsum[SCAN_NT / 32];
  const int i = blockIdx.x * SCAN_NT + threadIdx.x;
  const long long v = i < Q ? (long long)lens[i] : 0;
  long long incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  long long wb = 0;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) wb += wsum[w];
  if (i < Q) offsets[i] = bsum[blockIdx.x] + wb + incl - v;
}

__global__ void __launch_bounds__(256) gather_rows(const int32_t* ids, const double* scores, const int32_t* lens, int Q,
                                                   long long stride, const long long* offsets, int32_t* out_ids,
                                                   double* out_scores) {
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
    const int n = lens[q];
    const long long o = offsets[q];
    const int32_t* si = ids + (size_t)q * stride;
    const double* ss = scores + (size_t)q * stride;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      out_ids[o + j] = si[j];
      out_scores[o + j] = ss[j];
    }
  }
}

}  // namespace

extern "C" size_t qf_compact_workspace_bytes(int Q) {
  if (Q < 1) return 0;
  return sizeof(long long) * (size_t)((Q + SCAN_NT - 1) / SCAN_NT);
}

extern "C" int qf_compact_fused(const int32_t* ids, const double* scores, const int32_t* lens, int Q, long long stride,
                                long long* offsets, int32_t* out_ids, double* out_scores, void* ws, size_t ws_bytes,
                                void* stream) {
  QF_REQUIRE(Q >= 0 && stride >= 1, QF_E_INVALID, "qf_compact_fused: bad sizes");
  if (Q == 0) return QF_OK;
  QF_REQUIRE(ids && scores && lens && offsets && out_ids && out_scores && ws, QF_E_INVALID,
             "qf_compact_fused: null pointer argument");
  QF_REQUIRE(ws_bytes >= qf_compact_workspace_bytes(Q), QF_E_WORKSPACE, "qf_compact_fused: workspace too small");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long* bsum = reinterpret_cast<long long*>(ws);
  const int nb = (Q + SCAN_NT - 1) / SCAN_NT;
  block_sums<<<nb, SCAN_NT, 0, st>>>(lens, Q, bsum);
  scan_block_sums<<<1, SCAN_NT, 0, st>>>(bsum, nb, offsets, Q);
  finish_offsets<<<nb, SCAN_NT, 0, st>>>(lens, Q, bsum, offsets);
  const int sms = qf::sm_count();
  const long long g = (long long)sms * 8;
  gather_rows<<<(unsigned)(Q < g ? Q : g), 256, 0, st>>>(ids, scores, lens, Q, stride, offsets, out_ids, out_scores);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}
