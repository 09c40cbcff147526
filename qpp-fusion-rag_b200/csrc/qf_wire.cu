// qf_wire: bit-packed host->device wire format of the run postings.
//
// The end-to-end step (bench.py `e2e`, pipeline.FusionStepGraph(wire=...)) is bound by the PCIe upload of the
// [Q][R][K] (doc id int32, score int32 micro-units) pairs -- 8 bytes per posting, 138 MB for the NQ shape at ~52 GB/s
//  -- while the kernels need a tenth of that time.  Dictionary-encoded doc ids of a 2.68 M
// document collection need 22 bits and min-max normalised 6-decimal scores 20 bits, so the host packs every posting
// into id_bits + sc_bits bits (42 -> 5.25 bytes, 90.6 MB per step) and one HBM-bound kernel expands them into the
// int32 arrays the QPP and fusion kernels read.  Lossless: ids and micro-unit scores come back bit for bit.
//
// Layout: posting i (row-major position in [Q][R][K]) occupies bits [i*W, (i+1)*W) of a little-endian stream of
// uint64 words, W = id_bits + sc_bits <= 64; low id_bits = doc id, next sc_bits = micro - sc_base.  Positions beyond
// lens[q][r] are zero-filled.  Blocks of 64 postings are W words long, so host threads pack independent blocks.
// Replaces nothing in the reference (its run files are read from disk, src/fusion.py:46-76); it is the transport
// between the packer (qf_pack_weighted / packing.pack_runs) and the device.
#include "qf_common.cuh"

#include <thread>
#include <vector>

namespace {

using u32 = unsigned int;
using u64 = unsigned long long;

__device__ __forceinline__ u64 take_bits(const u64* __restrict__ words, long long bit, int W) {
  const long long wi = bit >> 6;
  const int sh = (int)(bit & 63);
  u64 v = __ldg(words + wi) >> sh;
  if (sh + W > 64) v |= __ldg(words + wi + 1) << (64 - sh);
  return W == 64 ? v : (v & ((1ull << W) - 1ull));
}

// 4 consecutive postings per thread: 16-byte stores into both output arrays.
__global__ void __launch_bounds__(256) wire_unpack_kernel(const u64* __restrict__ words, long long n, int id_bits,
                                                          int sc_bits, int sc_base, int32_t* __restrict__ ids,
                                                          int32_t* __restrict__ micro) {
  const int W = id_bits + sc_bits;
  const u64 id_mask = id_bits == 64 ? ~0ull : ((1ull << id_bits) - 1ull);
  const long long n4 = n >> 2;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += (long long)gridDim.x * blockDim.x) {
    const long long i0 = g << 2;
    int id[4], sc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u64 v = take_bits(words, (i0 + j) * W, W);
      id[j] = (int)(v & id_mask);
      sc[j] = (int)((u32)(v >> id_bits) + (u32)sc_base);
    }
    *reinterpret_cast<int4*>(ids + i0) = make_int4(id[0], id[1], id[2], id[3]);
    if (micro) *reinterpret_cast<int4*>(micro + i0) = make_int4(sc[0], sc[1], sc[2], sc[3]);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const u64 v = take_bits(words, i * W, W);
    ids[i] = (int)(v & id_mask);
    if (micro) micro[i] = (int)((u32)(v >> id_bits) + (u32)sc_base);
  }
}

// Frame-of-reference blocks: scores of a run list are sorted, so 32 consecutive ones span a small range.  Block b =
// list * ceil(K/32) + i/32 stores base[b] = its smallest score and 32 x w-bit offsets (w = bits(max - min), 0 when the
// block is constant or empty) as w uint32 words at payload[off[b] .. off[b+1]).  Thread = 4 consecutive postings of one
// block (K % 4 == 0) or one posting (any K).
template <int VEC>
__global__ void __launch_bounds__(256) wire_for_unpack_kernel(const int32_t* __restrict__ base, const u32* __restrict__ off,
                                                              const u32* __restrict__ payload, long long lists, int K,
                                                              int bpl, int32_t* __restrict__ micro) {
  const long long n = lists * K, ng = n / VEC;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (long long)gridDim.x * blockDim.x) {
    const long long p0 = g * VEC;
    const long long l = p0 / K;
    const int i0 = (int)(p0 - l * K);
    const long long b = l * bpl + (i0 >> 5);
    const u32 o = __ldg(off + b), w = __ldg(off + b + 1) - o;
    const u32 bs = (u32)__ldg(base + b);
    const u32 mask = w >= 32 ? 0xFFFFFFFFu : ((1u << w) - 1u);
    int out[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      u32 v = 0;
      if (w) {
        const u32 bit = (u32)((i0 + j) & 31) * w;
        const u32 wi = bit >> 5, sh = bit & 31;
        v = __ldg(payload + o + wi) >> sh;
        if (sh + w > 32) v |= __ldg(payload + o + wi + 1) << (32 - sh);
        v &= mask;
      }
      out[j] = (int)(bs + v);
    }
    if (VEC == 4) *reinterpret_cast<int4*>(micro + p0) = make_int4(out[0], out[VEC > 1 ? 1 : 0], out[VEC > 2 ? 2 : 0], out[VEC > 3 ? 3 : 0]);
    else micro[p0] = out[0];
  }
}

int bits_for(u64 span) {        // bits needed for values 0..span
  int b = 1;
  while (b < 64 && (span >> b)) ++b;
  return b;
}

template <typename F>
void parallel_rows(long long rows, int n_threads, F fn) {
  if (n_threads < 1) n_threads = 1;
  if (rows < 4096 || n_threads == 1) {
    fn(0, rows, 0);
    return;
  }
  std::vector<std::thread> th;
  const long long per = (rows + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const long long a = t * per, b = a + per < rows ? a + per : rows;
    if (a >= b) break;
    th.emplace_back(fn, a, b, t);
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" size_t qf_wire_words(long long n, int id_bits, int sc_bits) {
  if (n < 0 || id_bits < 1 || sc_bits < 0 || id_bits > 31 || sc_bits > 32) return 0;    // sc_bits == 0: doc ids only
  const long long blocks = (n + 63) / 64;
  return (size_t)(blocks * (id_bits + sc_bits) + 1);       // + 1: the unpack kernel may read one word ahead
}

extern "C" int qf_wire_plan(const int32_t* doc_ids, const int32_t* micro, const int32_t* lens, int Q, int R, int K,
                            int n_threads, int* id_bits, int* sc_bits, int32_t* sc_base) {
  QF_REQUIRE(doc_ids && micro && lens && id_bits && sc_bits && sc_base && Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID,
             "qf_wire_plan: bad argument");
  const long long lists = (long long)Q * R;
  if (n_threads < 1) n_threads = 1;
  std::vector<long long> mx_id(n_threads, 0), mn_sc(n_threads, INT32_MAX), mx_sc(n_threads, INT32_MIN);
  std::vector<int> bad(n_threads, 0);
  parallel_rows(lists, n_threads, [&](long long a, long long b, int t) {
    long long mi = 0, lo = INT32_MAX, hi = INT32_MIN;
    int bd = 0;
    for (long long l = a; l < b; ++l) {
      const int len = lens[l];
      if (len < 0 || len > K) { bd = 1; continue; }
      const int32_t* id = doc_ids + l * K;
      const int32_t* sc = micro + l * K;
      for (int i = 0; i < len; ++i) {
        if (id[i] < 0) bd = 2;
        mi = id[i] > mi ? id[i] : mi;
        lo = sc[i] < lo ? sc[i] : lo;
        hi = sc[i] > hi ? sc[i] : hi;
      }
    }
    mx_id[t] = mi; mn_sc[t] = lo; mx_sc[t] = hi; bad[t] = bd;
  });
  long long mi = 0, lo = INT32_MAX, hi = INT32_MIN;
  for (int t = 0; t < n_threads; ++t) {
    QF_REQUIRE(bad[t] != 1, QF_E_INVALID, "qf_wire_plan: a list length is outside [0, K]");
    QF_REQUIRE(bad[t] != 2, QF_E_UNSUPPORTED, "qf_wire_plan: negative doc id inside a list (the wire format carries "
               "dictionary codes >= 0)");
    mi = mx_id[t] > mi ? mx_id[t] : mi;
    lo = mn_sc[t] < lo ? mn_sc[t] : lo;
    hi = mx_sc[t] > hi ? mx_sc[t] : hi;
  }
  if (lo > hi) lo = hi = 0;                     // no postings at all
  *id_bits = bits_for((u64)mi);
  *sc_bits = bits_for((u64)(hi - lo));
  *sc_base = (int32_t)lo;
  return QF_OK;
}

extern "C" int qf_wire_pack(const int32_t* doc_ids, const int32_t* micro, const int32_t* lens, int Q, int R, int K,
                            int id_bits, int sc_bits, int32_t sc_base, uint64_t* out, size_t out_words, int n_threads) {
  QF_REQUIRE(doc_ids && (micro || sc_bits == 0) && lens && out && Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID,
             "qf_wire_pack: bad argument");
  const long long n = (long long)Q * R * K;
  const size_t need = qf_wire_words(n, id_bits, sc_bits);
  QF_REQUIRE(need != 0, QF_E_INVALID, "qf_wire_pack: id_bits=%d / sc_bits=%d outside [1,31] / [1,32]", id_bits, sc_bits);
  QF_REQUIRE(out_words >= need, QF_E_WORKSPACE, "qf_wire_pack: output holds %zu words, %zu needed (qf_wire_words)",
             out_words, need);
  const int W = id_bits + sc_bits;
  const long long blocks = (n + 63) / 64;
  if (n_threads < 1) n_threads = 1;
  std::vector<int> bad(n_threads, 0);
  const u64 id_lim = 1ull << id_bits, sc_lim = 1ull << sc_bits;
  parallel_rows(blocks, n_threads, [&](long long a, long long b, int t) {
    int bd = 0;
    for (long long blk = a; blk < b; ++blk) {
      u64* o = reinterpret_cast<u64*>(out) + blk * W;
      u64 acc = 0;
      int fill = 0;
      const long long p0 = blk * 64, p1 = p0 + 64 < n ? p0 + 64 : n;
      int wrote = 0;
      for (long long p = p0; p < p1; ++p) {
        const long long l = p / K;
        const int i = (int)(p - l * K);
        u64 v = 0;
        if (i < lens[l]) {
          const u64 id = (u64)(u32)doc_ids[p];
          u64 sc = 0;
          if (sc_bits) {
            sc = (u64)((long long)micro[p] - (long long)sc_base);
            if ((long long)micro[p] < (long long)sc_base || sc >= sc_lim) bd = 1;
          }
          if (doc_ids[p] < 0 || id >= id_lim) bd = 1;
          v = id | (sc << id_bits);
        }
        acc |= v << fill;
        if (fill + W >= 64) {
          o[wrote++] = acc;
          acc = fill ? (v >> (64 - fill)) : 0ull;
          fill = fill + W - 64;
        } else {
          fill += W;
        }
      }
      if (fill > 0 && wrote < W) o[wrote++] = acc;
      while (wrote < W) o[wrote++] = 0ull;
    }
    bad[t] = bd;
  });
  reinterpret_cast<u64*>(out)[blocks * W] = 0ull;
  for (int t = 0; t < n_threads; ++t)
    QF_REQUIRE(!bad[t], QF_E_INVALID, "qf_wire_pack: a doc id or score does not fit id_bits=%d / sc_bits=%d with "
               "sc_base=%d (call qf_wire_plan)", id_bits, sc_bits, (int)sc_base);
  return QF_OK;
}

extern "C" int qf_wire_unpack(const uint64_t* wire, long long n, int id_bits, int sc_bits, int32_t sc_base,
                              int32_t* doc_ids, int32_t* micro, void* stream) {
  QF_REQUIRE(n >= 0 && qf_wire_words(n, id_bits, sc_bits) != 0, QF_E_INVALID,
             "qf_wire_unpack: bad sizes n=%lld id_bits=%d sc_bits=%d", n, id_bits, sc_bits);
  if (n == 0) return QF_OK;
  QF_REQUIRE(wire && doc_ids && (micro || sc_bits == 0), QF_E_INVALID, "qf_wire_unpack: null pointer argument");
  QF_REQUIRE(((reinterpret_cast<uintptr_t>(doc_ids) | reinterpret_cast<uintptr_t>(micro)) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(wire) & 7) == 0, QF_E_INVALID, "qf_wire_unpack: outputs must be 16-byte "
             "aligned, the wire 8-byte aligned");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const long long groups = (n >> 2) + 1;
  long long grid = (groups + 255) / 256;
  const long long cap = (long long)qf::sm_count() * 16;
  if (grid > cap) grid = cap;
  wire_unpack_kernel<<<(unsigned)grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const u64*>(wire), n, id_bits, sc_bits, (int)sc_base, doc_ids, micro);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

extern "C" long long qf_wire_for_blocks(int Q, int R, int K) {
  if (Q < 0 || R < 1 || K < 1) return -1;
  return (long long)Q * R * ((K + 31) / 32);
}

extern "C" int qf_wire_for_pack(const int32_t* micro, const int32_t* lens, int Q, int R, int K, int n_threads,
                                int32_t* base, uint32_t* off, uint32_t* payload, size_t payload_cap_words,
                                size_t* words_used) {
  QF_REQUIRE(micro && lens && base && off && payload && words_used && Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID,
             "qf_wire_for_pack: bad argument");
  const long long lists = (long long)Q * R;
  const int bpl = (K + 31) / 32;
  const long long nblocks = lists * bpl;
  QF_REQUIRE(payload_cap_words >= (size_t)nblocks * 32 + 1, QF_E_WORKSPACE, "qf_wire_for_pack: payload holds %zu words, "
             "the worst case needs %lld (32 per block of 32 postings, + 1)", payload_cap_words, nblocks * 32 + 1);
  if (n_threads < 1) n_threads = 1;
  std::vector<int> bad(n_threads, 0);
  // pass 1: per block min and width (off[b + 1] holds the width until the prefix sum)
  off[0] = 0;
  parallel_rows(lists, n_threads, [&](long long a, long long b, int t) {
    for (long long l = a; l < b; ++l) {
      const int len = lens[l];
      if (len < 0 || len > K) { bad[t] = 1; continue; }
      const int32_t* sc = micro + l * K;
      for (int j = 0; j < bpl; ++j) {
        const int i0 = j * 32, i1 = i0 + 32 < len ? i0 + 32 : len;
        long long lo = 0, hi = 0;
        if (i0 < i1) {
          lo = hi = sc[i0];
          for (int i = i0 + 1; i < i1; ++i) { lo = sc[i] < lo ? sc[i] : lo; hi = sc[i] > hi ? sc[i] : hi; }
        }
        base[l * bpl + j] = (int32_t)lo;
        const u64 span = (u64)(hi - lo);
        off[l * bpl + j + 1] = span ? (uint32_t)bits_for(span) : 0u;
      }
    }
  });
  for (int t = 0; t < n_threads; ++t) QF_REQUIRE(!bad[t], QF_E_INVALID, "qf_wire_for_pack: a list length is outside [0, K]");
  u64 run = 0;
  for (long long b = 0; b < nblocks; ++b) {       // serial prefix sum: one add per 32 postings
    const u32 w = off[b + 1];
    off[b] = (uint32_t)run;
    run += w;
  }
  QF_REQUIRE(run < (1ull << 32), QF_E_UNSUPPORTED, "qf_wire_for_pack: payload exceeds 2^32 words");
  off[nblocks] = (uint32_t)run;
  // pass 2: payloads (blocks are word aligned: independent)
  parallel_rows(lists, n_threads, [&](long long a, long long b, int) {
    for (long long l = a; l < b; ++l) {
      const int len = lens[l];
      const int32_t* sc = micro + l * K;
      for (int j = 0; j < bpl; ++j) {
        const long long blk = l * bpl + j;
        const u32 w = off[blk + 1] - off[blk];
        if (!w) continue;
        uint32_t* o = payload + off[blk];
        const long long bs = base[blk];
        u64 acc = 0;
        int fill = 0, wrote = 0;
        for (int k = 0; k < 32; ++k) {
          const int i = j * 32 + k;
          const u64 v = i < len ? (u64)((long long)sc[i] - bs) : 0ull;
          acc |= v << fill;
          fill += (int)w;
          while (fill >= 32) {
            o[wrote++] = (uint32_t)acc;
            acc >>= 32;
            fill -= 32;
          }
        }
      }
    }
  });
  payload[run] = 0u;                               // the unpack kernel may read one word ahead
  *words_used = (size_t)run + 1;
  return QF_OK;
}

extern "C" int qf_wire_for_unpack(const int32_t* base, const uint32_t* off, const uint32_t* payload, int Q, int R, int K,
                                  int32_t* micro, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID, "qf_wire_for_unpack: bad sizes Q=%d R=%d K=%d", Q, R, K);
  if (Q == 0) return QF_OK;
  QF_REQUIRE(base && off && payload && micro, QF_E_INVALID, "qf_wire_for_unpack: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const long long lists = (long long)Q * R;
  const int bpl = (K + 31) / 32;
  const bool vec = (K % 4 == 0) && (reinterpret_cast<uintptr_t>(micro) & 15) == 0;
  const long long groups = lists * K / (vec ? 4 : 1);
  long long grid = (groups + 255) / 256;
  const long long cap = (long long)qf::sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (vec) wire_for_unpack_kernel<4><<<(unsigned)grid, 256, 0, st>>>(base, off, payload, lists, K, bpl, micro);
  else wire_for_unpack_kernel<1><<<(unsigned)grid, 256, 0, st>>>(base, off, payload, lists, K, bpl, micro);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}
