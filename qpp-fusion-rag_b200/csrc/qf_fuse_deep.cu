// qf_fuse_deep: rank fusion of run lists too deep for the shared-memory kernels (R*K above qf_fuse_max_postings(),
// e.g. 6 retrievers x TREC depth 1000 with float64 scores, or 20 runs x depth 2000): the reference fuses arbitrary
// depth (src/fusion.py:266-278, 309-327, 361-377; ranx.fuse :182,201,226), so this path must exist even though it is
// not the fast one.  Same semantics and outputs as qf_fuse (include/qppfusion.h), bit for bit.
//
// One persistent CTA per query; the per-query state lives in a caller-supplied device workspace (L2-resident at these
// sizes: ~50 B per table slot) instead of shared memory.  Postings are named by their offset `off = r*K + i` inside
// the query's [R][K] block, which is monotone in the reference's visiting order (r ascending, then row order):
//   P1  join: open addressing over table[H] (H = power of two >= 2*R*K) holding the offset of the posting that
//       claimed the slot; one atomicCAS per probe step, doc ids compared through the offset (so doc id -1 is an
//       ordinary id).  Slots only go from EMPTY to claimed, hence all postings of one doc end in the same slot.
//   P3  accumulation list by list (barrier between lists): inside a list doc ids are distinct, so the read-modify-
//       write acc = fl(acc + v) is race free and happens in the reference's order; a doc id repeated inside a list
//       is detected by a write/verify tag and sends the query to an exact sequential path (one thread).
//   P2  first-seen order by a block scan over the postings.
//   P5  ranking: stable LSD radix sort (8-bit digits, 8 passes over a 64-bit descending key, identity passes
//       skipped) of (key, first-seen index) pairs, starting from first-seen order -- ties keep it.
#include "qf_fuse_shared.cuh"

namespace {

using namespace qf;
using u16 = unsigned short;
using u32 = unsigned int;
using u64 = unsigned long long;

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr u32 EMPTY = 0xFFFFFFFFu;
constexpr int DEEP_MAX_POSTINGS = 65535;     // out_counts / out_pos are 16-bit

struct DeepMisc {
  int start[FUSE_MAX_R + 1];
  int len[FUSE_MAX_R];
  double w[FUSE_MAX_R], mn[FUSE_MAX_R], den[FUSE_MAX_R];
  int status, flag_dup, single, n_u;
  int warp_tot[NW];
  u64 red_min[NW], red_max[NW];
};

struct DeepWs {          // one CTA's slice of the workspace
  u32* table;            // [H]  offset of the claiming posting
  u32* tag;              // [H]  write/verify tag of the accumulation pass
  u32* first;            // [H]  offset of the doc's first-seen posting
  u32* cnt;              // [H]  postings merged (MNZ count)
  double* acc;           // [H]
  u32* slotof;           // [NP] slot of posting `off`
  u32* uslot;            // [NP] slot of the u-th doc in first-seen order
  u64* keyA;             // [NP] sort keys (double-buffered)
  u64* keyB;
  u32* idxA;             // [NP] first-seen indices (double-buffered)
  u32* idxB;
};

__host__ __device__ inline size_t a256(size_t x) { return (x + 255) & ~size_t(255); }
__host__ __device__ inline u32 table_size(long long n_max) {
  u32 h = 1024;
  while ((long long)h < 2 * n_max) h <<= 1;
  return h;
}
__host__ __device__ inline size_t n_pad_of(long long n_max) { return (size_t)((n_max + 31) & ~31ll); }
__host__ __device__ inline size_t per_cta_bytes(long long n_max) {
  const size_t H = table_size(n_max), NP = n_pad_of(n_max);
  return a256(4 * H) * 4 + a256(8 * H) + a256(4 * NP) * 2 + a256(8 * NP) * 2 + a256(4 * NP) * 2;
}
__device__ inline DeepWs carve(unsigned char* base, long long n_max) {
  const size_t H = table_size(n_max), NP = n_pad_of(n_max);
  DeepWs w;
  unsigned char* p = base;
  w.table = reinterpret_cast<u32*>(p); p += a256(4 * H);
  w.tag = reinterpret_cast<u32*>(p); p += a256(4 * H);
  w.first = reinterpret_cast<u32*>(p); p += a256(4 * H);
  w.cnt = reinterpret_cast<u32*>(p); p += a256(4 * H);
  w.acc = reinterpret_cast<double*>(p); p += a256(8 * H);
  w.slotof = reinterpret_cast<u32*>(p); p += a256(4 * NP);
  w.uslot = reinterpret_cast<u32*>(p); p += a256(4 * NP);
  w.keyA = reinterpret_cast<u64*>(p); p += a256(8 * NP);
  w.keyB = reinterpret_cast<u64*>(p); p += a256(8 * NP);
  w.idxA = reinterpret_cast<u32*>(p); p += a256(4 * NP);
  w.idxB = reinterpret_cast<u32*>(p);
  return w;
}

// One stable LSD pass over n_pad (multiple of 32) (key, idx) pairs in global memory.  Each warp owns a contiguous
// range of 32-element rounds; ranking inside a round by __match_any_sync, per-warp digit counters whist[digit][warp]
// in shared memory (only a digit group's lowest lane touches its counter: no atomics).  The digits are recomputed in
// the scatter walk, which visits the rounds in the same order.  Returns false (nothing moved) for a constant digit.
__device__ bool deep_radix_pass(const u64* keyIn, u64* keyOut, const u32* idxIn, u32* idxOut, int n_pad, int shift,
                                u32* whist, DeepMisc& m) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rounds_total = n_pad >> 5;
  const int rpw = (rounds_total + NW - 1) / NW;
  const int r0 = warp * rpw, r1 = min(rounds_total, r0 + rpw);
  for (int i = tid; i < 256 * NW; i += NT) whist[i] = 0u;
  if (tid == 0) m.single = 0;
  __syncthreads();
  for (int rd = r0; rd < r1; ++rd) {
    const u32 d = (u32)(keyIn[(rd << 5) + lane] >> shift) & 255u;
    const u32 peers = __match_any_sync(0xffffffffu, d);
    if ((peers & ((1u << lane) - 1u)) == 0) whist[d * NW + warp] += (u32)__popc(peers);
    __syncwarp();
  }
  __syncthreads();
  {                                   // digit-major exclusive scan: thread d < 256 owns digit d
    u32 row[NW];
    int tot = 0;
    if (tid < 256) {
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        row[w] = (u32)tot;
        tot += (int)whist[tid * NW + w];
      }
      if (tot == n_pad) m.single = 1;
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (tid < 256 && lane == 31) m.warp_tot[warp] = incl;
    __syncthreads();
    if (tid < 256) {
      int excl = incl - tot;
      for (int w = 0; w < warp; ++w) excl += m.warp_tot[w];
#pragma unroll
      for (int w = 0; w < NW; ++w) whist[tid * NW + w] = row[w] + (u32)excl;
    }
    __syncthreads();
  }
  if (m.single) {
    __syncthreads();
    return false;
  }
  for (int rd = r0; rd < r1; ++rd) {
    const int e = (rd << 5) + lane;
    const u64 key = keyIn[e];
    const u32 d = (u32)(key >> shift) & 255u;
    const u32 peers = __match_any_sync(0xffffffffu, d);
    const u32 lt = peers & ((1u << lane) - 1u);
    u32 base = 0;
    if (lt == 0) {
      base = whist[d * NW + warp];
      whist[d * NW + warp] = base + (u32)__popc(peers);
    }
    __syncwarp();
    base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
    const u32 pos = base + (u32)__popc(lt);
    keyOut[pos] = key;
    idxOut[pos] = idxIn[e];
  }
  __syncthreads();
  return true;
}

template <int ST>
__global__ void __launch_bounds__(NT) fuse_deep_kernel(const FuseParams p, unsigned char* ws, size_t cta_bytes) {
  using raw_t = typename ScoreT<ST>::type;
  __shared__ u32 whist[256 * NW];
  __shared__ DeepMisc m;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, K = p.K;
  const long long n_max = (long long)R * K;
  const size_t out_stride = (size_t)n_max;
  const u32 H = table_size(n_max), HM = H - 1;
  const DeepWs w = carve(ws + (size_t)blockIdx.x * cta_bytes, n_max);
  const bool is_rrf = (p.method == QF_METHOD_RRF);
  const bool is_mnz = (p.method == QF_METHOD_MNZ);
  const bool skip_zero = (p.flags & QF_FLAG_MNZ_SKIP_ZERO) != 0;
  const bool normed = (p.norm != QF_NORM_NONE) && !is_rrf;

  for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
    const int32_t* gid = p.doc_ids + (size_t)q * n_max;
    const raw_t* gsc = reinterpret_cast<const raw_t*>(p.scores) + (size_t)q * n_max;
    const int32_t* grk = (is_rrf && p.ranks) ? p.ranks + (size_t)q * n_max : nullptr;
    // ---------------------------------------------------------------- header, table reset
    if (tid == 0) {
      int acc = 0, bad = 0;
      for (int r = 0; r < R; ++r) {
        m.start[r] = acc;
        const int len = p.lens[(size_t)q * R + r];
        if (len < 0 || len > K) bad = 1;
        m.len[r] = bad ? 0 : len;
        acc += m.len[r];
        m.w[r] = p.weights ? p.weights[(size_t)q * R + r] : 1.0;
      }
      m.start[R] = acc;
      m.status = bad ? QF_QSTAT_BAD_LEN : 0;
      m.flag_dup = 0;
    }
    for (u32 i = tid; i < H; i += NT) w.table[i] = EMPTY;     // the per-slot state is initialised by the claiming posting
    __syncthreads();
    const int n = m.start[R];
    if ((m.status & QF_QSTAT_BAD_LEN) || n == 0) {
      if (tid == 0) {
        p.out_lens[q] = 0;
        if (p.out_status) p.out_status[q] = m.status;
      }
      __syncthreads();
      continue;
    }
    // ---------------------------------------------------------------- per-list normalisation constants, validation
    if (normed) {
      for (int r = warp; r < R; r += NW) {
        const raw_t* gs = gsc + (size_t)r * K;
        const int len = m.len[r];
        if (p.norm == QF_NORM_ZSCORE) {        // sequential float64 sums in list order (include/qppfusion.h)
          if (lane == 0 && len > 0) {
            double c = 0.0;
            for (int i = 0; i < len; ++i) c = __dadd_rn(c, raw_to_f64(gs[i]));
            const double mean = __ddiv_rn(c, (double)len);
            double ssd = 0.0;
            for (int i = 0; i < len; ++i) {
              const double d = __dsub_rn(raw_to_f64(gs[i]), mean);
              ssd = __dadd_rn(ssd, __dmul_rn(d, d));
            }
            const double sd = __dsqrt_rn(__ddiv_rn(ssd, (double)len));
            m.mn[r] = mean;
            m.den[r] = sd > 1e-9 ? sd : 1e-9;
          }
        } else {
          double mn = INFINITY, mx = -INFINITY;
          for (int i = lane; i < len; i += 32) {
            const double v = raw_to_f64(gs[i]);
            mn = fmin(mn, v);
            mx = fmax(mx, v);
          }
          mn = warp_min(mn);
          mx = warp_max(mx);
          if (lane == 0) {
            m.mn[r] = mn;
            const double d = __dsub_rn(mx, mn);
            m.den[r] = d > 1e-9 ? d : 1e-9;       // ranx: max(max - min, 1e-9)
          }
        }
      }
    }
    {
      int bad = 0;
      for (int r = 0; r < R; ++r) {
        const int len = m.len[r];
        if (p.flags & QF_FLAG_REQUIRE_SORTED)
          for (int i = tid + 1; i < len; i += NT)
            if (raw_to_f64(gsc[(size_t)r * K + i - 1]) < raw_to_f64(gsc[(size_t)r * K + i])) bad = 1;
        if (grk)
          for (int i = tid + 1; i < len; i += NT)
            if (grk[(size_t)r * K + i - 1] > grk[(size_t)r * K + i]) bad = 1;
      }
      if (bad) atomicOr(&m.status, QF_QSTAT_UNSORTED);
    }
    // ---------------------------------------------------------------- P1: join
    for (int r = 0; r < R; ++r) {
      const int len = m.len[r];
      for (int i = tid; i < len; i += NT) {
        const u32 off = (u32)r * (u32)K + (u32)i;
        const int32_t id = gid[off];
        u32 h = ((u32)id * 2654435761u) >> 7 & HM;
        while (true) {
          const u32 o = atomicCAS(&w.table[h], EMPTY, off);
          if (o == EMPTY) {                      // claimed: this posting initialises the slot (read after the barrier)
            w.first[h] = EMPTY;
            w.cnt[h] = 0u;
            w.acc[h] = 0.0;
            break;
          }
          if (gid[o] == id) break;
          h = (h + 1) & HM;
        }
        w.slotof[off] = h;
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- P3: ordered accumulation
    auto contribution = [&](u32 off, int r, int i, bool& counts) -> double {
      counts = true;
      if (is_rrf) {
        const double rank = grk ? (double)grk[off] : (double)(i + 1);
        return __ddiv_rn(m.w[r], __dadd_rn(p.rrf_k, rank));
      }
      double s = raw_to_f64(gsc[off]);
      if (normed) s = __ddiv_rn(__dsub_rn(s, m.mn[r]), m.den[r]);
      if (skip_zero && s == 0.0) counts = false;
      return __dmul_rn(m.w[r], s);                 // w == 1.0 when unweighted: exact
    };
    for (int r = 0; r < R; ++r) {
      const int len = m.len[r];
      for (int i = tid; i < len; i += NT) {
        const u32 off = (u32)r * (u32)K + (u32)i;
        w.tag[w.slotof[off]] = off;
      }
      __syncthreads();
      for (int i = tid; i < len; i += NT) {
        const u32 off = (u32)r * (u32)K + (u32)i;
        const u32 slot = w.slotof[off];
        if (w.tag[slot] != off) {
          m.flag_dup = 1;
        } else {
          bool counts;
          const double v = contribution(off, r, i, counts);
          if (w.first[slot] == EMPTY) w.first[slot] = off;
          w.acc[slot] = __dadd_rn(w.acc[slot], v);
          if (counts) w.cnt[slot] += 1u;
        }
      }
      __syncthreads();
    }
    if (m.flag_dup) {
      // exact sequential path: a doc id repeats inside one list (reference: every row accumulates and counts,
      // src/fusion.py:275,323-324)
      for (u32 i = tid; i < H; i += NT)
        if (w.table[i] != EMPTY) {
          w.first[i] = EMPTY;
          w.cnt[i] = 0u;
          w.acc[i] = 0.0;
        }
      __syncthreads();
      if (tid == 0) {
        for (int r = 0; r < R; ++r)
          for (int i = 0; i < m.len[r]; ++i) {
            const u32 off = (u32)r * (u32)K + (u32)i;
            const u32 slot = w.slotof[off];
            bool counts;
            const double v = contribution(off, r, i, counts);
            if (w.first[slot] == EMPTY) w.first[slot] = off;
            w.acc[slot] = __dadd_rn(w.acc[slot], v);
            if (counts) w.cnt[slot] += 1u;
          }
        m.status |= QF_QSTAT_DUPLICATES | ((p.flags & QF_FLAG_REQUIRE_SORTED) ? QF_QSTAT_UNSORTED : 0);
      }
      __syncthreads();
    }
    // ---------------------------------------------------------------- P2: first-seen order
    {
      const int ppt = (n + NT - 1) / NT;              // consecutive postings (in visiting order) per thread
      const int s_lo = min(n, tid * ppt), s_hi = min(n, s_lo + ppt);
      int r = 0;
      while (r + 1 < R && s_lo >= m.start[r + 1]) ++r;
      int cnt = 0;
      for (int seq = s_lo, rr = r; seq < s_hi; ++seq) {
        while (seq >= m.start[rr + 1]) ++rr;
        const u32 off = (u32)rr * (u32)K + (u32)(seq - m.start[rr]);
        cnt += (w.first[w.slotof[off]] == off) ? 1 : 0;
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) m.warp_tot[warp] = incl;
      __syncthreads();
      int u = incl - cnt;
      for (int ww = 0; ww < warp; ++ww) u += m.warp_tot[ww];
      for (int seq = s_lo, rr = r; seq < s_hi; ++seq) {
        while (seq >= m.start[rr + 1]) ++rr;
        const u32 off = (u32)rr * (u32)K + (u32)(seq - m.start[rr]);
        const u32 slot = w.slotof[off];
        if (w.first[slot] == off) w.uslot[u++] = slot;
      }
      if (tid == NT - 1) m.n_u = u;
    }
    __syncthreads();
    const int n_u = m.n_u;
    // ---------------------------------------------------------------- P4: final scores and key range
    u64 omin = ~0ull, omax = 0ull;
    for (int u = tid; u < n_u; u += NT) {
      const u32 slot = w.uslot[u];
      double fs = w.acc[slot];
      if (is_mnz) fs = __dmul_rn(fs, (double)w.cnt[slot]);
      fs = fs + 0.0;                                   // -0.0 -> +0.0 (Python: -0.0 == 0.0 is a tie)
      w.acc[slot] = fs;
      const u64 o = ord_f64(fs);
      omin = o < omin ? o : omin;
      omax = o > omax ? o : omax;
    }
    omin = warp_min(omin);
    omax = warp_max(omax);
    if (lane == 0) {
      m.red_min[warp] = omin;
      m.red_max[warp] = omax;
    }
    __syncthreads();
    omin = ~0ull;
    omax = 0ull;
    for (int ww = 0; ww < NW; ++ww) {
      omin = m.red_min[ww] < omin ? m.red_min[ww] : omin;
      omax = m.red_max[ww] > omax ? m.red_max[ww] : omax;
    }
    // ---------------------------------------------------------------- P5: ranking
    const u32* rank_idx = nullptr;            // nullptr: first-seen order is the output order
    if (!(p.flags & QF_FLAG_UNRANKED_OUTPUT) && n_u > 1 && omax != omin) {
      const int n_pad = (n_u + 31) & ~31;
      for (int u = tid; u < n_pad; u += NT) {
        w.keyA[u] = u < n_u ? omax - ord_f64(w.acc[w.uslot[u]]) : ~0ull;   // ascending key = descending score
        w.idxA[u] = (u32)u;                   // padding keeps the largest key and starts at the tail: stays there
      }
      __syncthreads();
      const int top = 64 - __clzll((long long)(omax - omin));               // bits that differ at all
      u64 *kin = w.keyA, *kout = w.keyB;
      u32 *iin = w.idxA, *iout = w.idxB;
      for (int shift = 0; shift < 64; shift += 8) {
        if (shift >= top) break;     // higher digits: 0 for every doc, 255 for the padding already at the tail
        if (deep_radix_pass(kin, kout, iin, iout, n_pad, shift, whist, m)) {
          u64* tk = kin; kin = kout; kout = tk;
          u32* ti = iin; iin = iout; iout = ti;
        }
      }
      rank_idx = iin;
    }
    // ---------------------------------------------------------------- output
    {
      int32_t* oid = p.out_ids + (size_t)q * out_stride;
      double* osc = p.out_scores + (size_t)q * out_stride;
      for (int j = tid; j < n_u; j += NT) {
        const u32 u = rank_idx ? rank_idx[j] : (u32)j;
        const u32 slot = w.uslot[u];
        oid[j] = gid[w.first[slot]];
        osc[j] = w.acc[slot];
        if (p.out_counts) {
          const u32 c = w.cnt[slot];
          p.out_counts[(size_t)q * out_stride + j] = (u16)(c > 0xFFFFu ? 0xFFFFu : c);
        }
        if (p.out_pos) p.out_pos[(size_t)q * out_stride + j] = (u16)u;
      }
      if (tid == 0) {
        p.out_lens[q] = n_u;
        if (p.out_status) p.out_status[q] = m.status;
      }
    }
    __syncthreads();
  }
}

int deep_grid(int Q, size_t ws_bytes, size_t cta_bytes) {
  long long grid = 2ll * qf::sm_count();
  if (grid > Q) grid = Q;
  const long long fit = (long long)(ws_bytes / cta_bytes);
  if (grid > fit) grid = fit;
  return (int)grid;
}

}  // namespace

extern "C" int qf_fuse_deep_max_postings(void) { return DEEP_MAX_POSTINGS; }

extern "C" size_t qf_fuse_deep_workspace_bytes(int Q, int R, int K) {
  if (Q <= 0 || R <= 0 || K <= 0 || (long long)R * K > DEEP_MAX_POSTINGS) return 0;
  long long ctas = 2ll * 148;
  if (ctas > Q) ctas = Q;
  return (size_t)ctas * per_cta_bytes((long long)R * K);
}

extern "C" int qf_fuse_deep(const int32_t* doc_ids, const void* scores, int score_type, const int32_t* ranks,
                            const int32_t* lens, const double* weights, int Q, int R, int K, int method, int norm,
                            double rrf_k, uint32_t flags, int32_t* out_ids, double* out_scores, uint16_t* out_counts,
                            uint16_t* out_pos, int32_t* out_lens, int32_t* out_status, void* ws, size_t ws_bytes,
                            void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID, "qf_fuse_deep: bad sizes Q=%d R=%d K=%d", Q, R, K);
  QF_REQUIRE(R <= FUSE_MAX_R, QF_E_UNSUPPORTED, "qf_fuse_deep: R=%d exceeds %d", R, FUSE_MAX_R);
  QF_REQUIRE((long long)R * K <= DEEP_MAX_POSTINGS, QF_E_UNSUPPORTED,
             "qf_fuse_deep: R*K=%lld exceeds %d postings per query (16-bit count / first-seen outputs)",
             (long long)R * K, DEEP_MAX_POSTINGS);
  QF_REQUIRE(method == QF_METHOD_SUM || method == QF_METHOD_MNZ || method == QF_METHOD_RRF, QF_E_INVALID,
             "qf_fuse_deep: unknown method %d", method);
  QF_REQUIRE(norm == QF_NORM_NONE || norm == QF_NORM_MINMAX || norm == QF_NORM_ZSCORE, QF_E_INVALID,
             "qf_fuse_deep: unknown norm %d", norm);
  QF_REQUIRE(score_type >= QF_SCORE_F64 && score_type <= QF_SCORE_MICRO, QF_E_INVALID,
             "qf_fuse_deep: unknown score_type %d", score_type);
  if (Q == 0) return QF_OK;
  QF_REQUIRE(doc_ids && scores && lens && out_ids && out_scores && out_lens, QF_E_INVALID,
             "qf_fuse_deep: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const size_t cta_bytes = per_cta_bytes((long long)R * K);
  QF_REQUIRE(ws != nullptr && (reinterpret_cast<uintptr_t>(ws) & 255) == 0, QF_E_WORKSPACE,
             "qf_fuse_deep: workspace must be a 256-byte aligned device buffer");
  const int grid = deep_grid(Q, ws_bytes, cta_bytes);
  QF_REQUIRE(grid >= 1, QF_E_WORKSPACE, "qf_fuse_deep: workspace of %zu bytes holds no query (one needs %zu; call "
             "qf_fuse_deep_workspace_bytes)", ws_bytes, cta_bytes);
  FuseParams p{doc_ids, scores, ranks, lens, weights, Q, R, K, method, norm, rrf_k, flags,
               out_ids, out_scores, out_counts, out_pos, out_lens, out_status, 0};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  qf::prof_mark(QF_PROF_FUSE_GENERAL, 0, st);
  switch (score_type) {
    case QF_SCORE_F64: fuse_deep_kernel<QF_SCORE_F64><<<grid, NT, 0, st>>>(p, wsb, cta_bytes); break;
    case QF_SCORE_F32: fuse_deep_kernel<QF_SCORE_F32><<<grid, NT, 0, st>>>(p, wsb, cta_bytes); break;
    default: fuse_deep_kernel<QF_SCORE_MICRO><<<grid, NT, 0, st>>>(p, wsb, cta_bytes); break;
  }
  qf::prof_mark(QF_PROF_FUSE_GENERAL, 1, st);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}
