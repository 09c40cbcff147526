// qf_fuse_stream: rank fusion without staging -- every posting is read from HBM exactly once, straight into the
// hash-accumulate step; several CTAs per SM (one query each) so one query's barriers hide behind another's work.
//
// Replaces ranx.fuse (reference src/fusion.py:182,201,226) and the weighted_* Python dict loops
// (src/fusion.py:266-278, 309-327, 361-377); semantics in include/qppfusion.h.  Per query:
//   A1 join, doc ids only: finding a doc's slot does not depend on the order of the postings, so every posting of
//      every list (16-byte loads, next chunk prefetched into registers, no barrier between lists) finds / claims its
//      doc's slot of a bucketised open-addressing table and records it in s_slot[r*K + i].  One 16-byte shared-memory
//      load fetches the 4 keys of the home group, a match needs no atomic, a miss claims the first empty key with one
//      predicated atomicCAS; a thread's four postings are resolved side by side in straight-line code.  Postings whose
//      home group is full or whose claim was lost are queued and resolved once per query on dense lanes (probe loop
//      uniform over the warp, per-doc odd stride).  The first non-empty list is inserted without probing through
//      per-group fill counters.
//   A2 accumulate, list by list (one barrier per list, no atomics): posting i of list r adds fl(w_r * s_i) to its
//      slot's float64 accumulator with a plain read-modify-write -- doc ids are distinct inside a list and lists are
//      separated by barriers, so the additions happen in exactly the reference's order ((w0*s0 + w1*s1) + ...), no
//      FMA contraction.  The posting that finds the slot's list mask empty is the doc's FIRST-SEEN posting; its
//      sequence number goes into the upper half of the meta word (tie-break key).  A doc id occurring twice inside a
//      list shows up as a deficit in the sum of the mask popcounts -> the query is deferred to the general kernel's
//      exact sequential path.
//   B  final score (MNZ multiply by popcount, -0.0 -> +0.0), min/max over the union.
//   C  ranking by ONE bucket pass instead of a multi-pass radix sort: a 64-bin histogram of a sample
//      gives a monotone piecewise-linear map score -> bucket (histogram equalisation, so skewed score
//      distributions still spread evenly over the buckets); shared-memory atomicAdd assigns a slot in the
//      bucket, a block scan turns counts into bases, and each doc's exact rank is its bucket base plus the number
//      of bucket mates that precede it under the exact comparison (score desc as float64, first-seen asc).
//      Buckets longer than LONGB (long runs of exactly tied scores) defer the query.
//   D  (id, score[, count, first-seen index]) written straight to the ranked position (bucket order makes the
//      stores nearly coalesced).
// Deferred queries (out_lens = FUSE_DEFERRED) are finished by qf_fuse.cu's general kernel in the same stream.
#include "qf_fuse_shared.cuh"

#include <stdlib.h>

namespace {

using namespace qf;
using u16 = unsigned short;
using u32 = unsigned int;
using u64 = unsigned long long;

constexpr u32 EMPTY = 0xFFFFFFFFu;
constexpr int SMAX_R = 16;       // list bit mask is the low half of the meta word
constexpr int CB = 64;           // coarse bins of the equalisation histogram
constexpr int MAXPROBE = 64;     // key groups (of 4) visited before the query is deferred
constexpr int LONGB = 160;       // longest bucket ranked by pairwise comparison

// Shared-memory accesses by 32-bit shared-window address (no generic->shared conversion per access).
__device__ __forceinline__ uint4 lds128_volatile(u32 sa) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
  return v;
}
// CAS executed only where `pred` holds; elsewhere the result is `dflt` (no branch around the atomic).
__device__ __forceinline__ u32 atoms_cas_if(bool pred, u32 sa, u32 cmp, u32 val, u32 dflt) {
  u32 old = dflt;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p atom.shared.cas.b32 %0, [%1], %2, %3;\n\t}"
               : "+r"(old) : "r"(sa), "r"(cmp), "r"(val), "r"((u32)pred) : "memory");
  return old;
}
// shared-memory add executed only where `pred` holds; elsewhere the result is `dflt`
__device__ __forceinline__ u32 atoms_add_if(bool pred, u32 sa, u32 val, u32 dflt) {
  u32 old = dflt;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
               : "+r"(old) : "r"(sa), "r"(val), "r"((u32)pred) : "memory");
  return old;
}
__device__ __forceinline__ u32 atoms_add(u32 sa, u32 val) {
  u32 old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(sa), "r"(val) : "memory");
  return old;
}
struct SMisc {
  int start[SMAX_R + 1];
  int len[SMAX_R];
  int next[SMAX_R];  // next non-empty list after r (R if none)
  int first;         // first non-empty list
  double w[SMAX_R], mn[SMAX_R], den[SMAX_R];
  double red_mn[16], red_mx[16];
  int red_cnt[16];
  int warp_tot[16];
  int status;
  int qn;           // queued (unresolved) postings of the join pass
  int ins_total;    // claimed keys = size of the union
  int pc;           // sum of the list-mask popcounts (== number of postings unless a list holds a doc id twice)
};

// Classify one group of 4 keys with a single select chain: 0..3 = position of `id`, 4..7 = 4 + position of the
// first empty key (no match), -1 = four foreign keys.
__device__ __forceinline__ int probe4(const uint4& k, u32 id) {
  int pos;
  asm("{\n\t.reg .pred p;\n\t"
      "mov.s32 %0, -1;\n\t"
      "setp.eq.u32 p, %4, 0xffffffff;\n\tselp.s32 %0, 7, %0, p;\n\t"
      "setp.eq.u32 p, %3, 0xffffffff;\n\tselp.s32 %0, 6, %0, p;\n\t"
      "setp.eq.u32 p, %2, 0xffffffff;\n\tselp.s32 %0, 5, %0, p;\n\t"
      "setp.eq.u32 p, %1, 0xffffffff;\n\tselp.s32 %0, 4, %0, p;\n\t"
      "setp.eq.u32 p, %4, %5;\n\tselp.s32 %0, 3, %0, p;\n\t"
      "setp.eq.u32 p, %3, %5;\n\tselp.s32 %0, 2, %0, p;\n\t"
      "setp.eq.u32 p, %2, %5;\n\tselp.s32 %0, 1, %0, p;\n\t"
      "setp.eq.u32 p, %1, %5;\n\tselp.s32 %0, 0, %0, p;\n\t"
      "}"
      : "=r"(pos)
      : "r"(k.x), "r"(k.y), "r"(k.z), "r"(k.w), "r"(id));
  return pos;
}

template <int G, int UCAP, int NB, int NMAX, int NT>
struct SLayout {
  static constexpr int H = 4 * G;                              // slots: G groups of 4 keys (one 16-byte load per probe)
  static constexpr int BW = NMAX / 32;                         // first-seen bitmap words
  static constexpr int CPT = NB / NT;                          // bucket counters scanned per thread
  static constexpr size_t a16(size_t x) { return (x + 15) & ~size_t(15); }
  static constexpr size_t off_keys = 0;
  static constexpr size_t off_meta = off_keys + 4 * H;
  static constexpr size_t off_acc = off_meta + 4 * H;
  static constexpr size_t off_cnt = off_acc + 8 * H;           // u16 bucket counters -> bases
  static constexpr size_t off_order = a16(off_cnt + 2 * NB);   // u16 order[UCAP]: slots in bucket order
  static constexpr size_t off_slot = off_cnt;                  // phase A only: u16 slot of posting r*K+i [NMAX]
  static constexpr size_t off_bcache = a16(off_order + 2 * UCAP);   // u16 bucket id per slot
  static constexpr size_t off_queue = off_bcache;                   // phase A only: u16 queue of unresolved postings [H]
  static constexpr size_t off_bitmap = a16(off_bcache + 2 * H);     // first-seen bitmap + word prefix (out_pos)
  static constexpr size_t off_coarse = a16(off_bitmap + 6 * BW);
  static constexpr size_t off_misc = a16(off_coarse + 4 * CB * 3);
  static constexpr size_t total = a16(off_misc + sizeof(SMisc));
  static_assert(NB % NT == 0 && CPT % 8 == 0 && NB < 65536, "bucket scan: CPT counters per thread, 8 per 16-byte load");
  static_assert(UCAP < H && NMAX % 32 == 0 && NMAX <= 65536 && UCAP < 65536 && H < 65536, "layout");
  static_assert(NT / 32 <= 16 && NT / 32 >= 2, "warps");
  static_assert(off_slot + 2 * NMAX <= off_bcache, "s_slot must fit into the bucket counter + order arrays");
};

__device__ __forceinline__ int warp_sum_int(int v) { return __reduce_add_sync(0xffffffffu, v); }

// MODE 0: CombSUM / CombMNZ (weighted or not) on raw scores -- no min-max norm, no zero-skipping, no order checks
//         (the weighted_* path and the benchmark step); MODE 1: every option decided at run time.
template <int G, int UCAP, int NB, int NMAX, int NT, int MINB, int ST, int MODE>
__global__ void __launch_bounds__(NT, MINB) fuse_stream_kernel(const FuseParams p) {
  using L = SLayout<G, UCAP, NB, NMAX, NT>;
  using raw_t = typename ScoreT<ST>::type;
  constexpr int H = L::H, NW = NT / 32, SPT = (H + NT - 1) / NT, CH = NT * 4, BW = L::BW, CPT = L::CPT, QCAP = H;

  extern __shared__ __align__(16) unsigned char smem[];
  u32* s_keys = reinterpret_cast<u32*>(smem + L::off_keys);
  u32* s_meta = reinterpret_cast<u32*>(smem + L::off_meta);
  double* s_acc = reinterpret_cast<double*>(smem + L::off_acc);
  u32* s_cnt32 = reinterpret_cast<u32*>(smem + L::off_cnt);
  u16* s_base16 = reinterpret_cast<u16*>(smem + L::off_cnt);
  u16* s_order = reinterpret_cast<u16*>(smem + L::off_order);
  u16* s_bcache = reinterpret_cast<u16*>(smem + L::off_bcache);
  u32* s_bitmap = reinterpret_cast<u32*>(smem + L::off_bitmap);
  u16* s_bpre = reinterpret_cast<u16*>(smem + L::off_bitmap + 4 * BW);
  u16* s_queue = reinterpret_cast<u16*>(smem + L::off_queue);   // phase A: postings still to be resolved
  u16* s_slot = reinterpret_cast<u16*>(smem + L::off_slot);     // phase A: table slot of every posting
  int* s_chist = reinterpret_cast<int*>(smem + L::off_coarse);
  float* s_ccum = reinterpret_cast<float*>(smem + L::off_coarse + 4 * CB);
  float* s_cw = reinterpret_cast<float*>(smem + L::off_coarse + 8 * CB);
  SMisc& m = *reinterpret_cast<SMisc*>(smem + L::off_misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, K = p.K;
  const size_t out_stride = (size_t)R * K;
  const u32 keys_sa = (u32)__cvta_generic_to_shared(s_keys);
  const u32 meta_sa = (u32)__cvta_generic_to_shared(s_meta);
  const bool is_rrf = MODE ? (p.method == QF_METHOD_RRF) : false;
  const bool is_mnz = (p.method == QF_METHOD_MNZ);
  const bool weighted = (p.weights != nullptr);
  const bool skip_zero = MODE ? (p.flags & QF_FLAG_MNZ_SKIP_ZERO) != 0 : false;
  const bool check_sorted = MODE ? (p.flags & QF_FLAG_REQUIRE_SORTED) != 0 : false;
  const bool unranked = (p.flags & QF_FLAG_UNRANKED_OUTPUT) != 0;
  const bool zscore = MODE ? ((p.norm == QF_NORM_ZSCORE) && !is_rrf) : false;
  const bool minmax = MODE ? ((p.norm != QF_NORM_NONE) && !is_rrf) : false;     // any per-list (sub, den) normalisation
  const bool need_scores = !is_rrf || check_sorted;
  const bool want_pos = (p.out_pos != nullptr);

  // warp 0 keeps the NEXT query's list lengths and weights in registers (loaded while the current one is fused)
  int pre_len = 0;
  double pre_w = 1.0;
  if (warp == 0 && lane < R && (int)blockIdx.x < p.Q) {
    pre_len = __ldg(p.lens + (size_t)blockIdx.x * R + lane);
    if (weighted) pre_w = __ldg(p.weights + (size_t)blockIdx.x * R + lane);
  }
  const u32 qn_sa = (u32)__cvta_generic_to_shared(&m.qn);

  for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
    // ------------------------------------------------------------ header + table reset
    // warp 0 works the header out in registers BEFORE the barrier that retires the previous query
    int h_l = 0, h_incl = 0;
    double h_w = 1.0;
    unsigned h_badm = 0u, h_nz = 0u;
    if (warp == 0) {
      const int len = pre_len;
      h_w = pre_w;
      {
        const long long qx = (long long)q + gridDim.x;
        if (lane < R && qx < p.Q) {
          pre_len = __ldg(p.lens + (size_t)qx * R + lane);
          if (weighted) pre_w = __ldg(p.weights + (size_t)qx * R + lane);
        }
      }
      h_badm = __ballot_sync(0xffffffffu, lane < R && (len < 0 || len > K));
      h_l = h_badm ? 0 : len;
      h_incl = h_l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, h_incl, o);
        if (lane >= o) h_incl += t;
      }
      h_nz = __ballot_sync(0xffffffffu, lane < R && h_l > 0);
    }
    __syncthreads();                       // previous query's shared state is dead
    if (warp == 0) {
      if (lane < R) {
        m.start[lane] = h_incl - h_l;
        m.len[lane] = h_l;
        m.w[lane] = h_w;
        const unsigned rest = lane < 31 ? h_nz >> (lane + 1) : 0u;
        m.next[lane] = rest ? lane + __ffs(rest) : R;
      }
      if (lane == R - 1) m.start[R] = h_incl;
      if (lane == 0) {
        m.first = h_nz ? __ffs(h_nz) - 1 : R;
        m.status = h_badm ? QF_QSTAT_BAD_LEN : 0;
        m.qn = 0;
        m.ins_total = 0;
        m.pc = 0;
      }
    }
    {
      uint4* k4 = reinterpret_cast<uint4*>(s_keys);
      uint4* m4 = reinterpret_cast<uint4*>(s_meta);
      for (int i = tid; i < H / 4; i += NT) {
        k4[i] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
        m4[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (want_pos)
        for (int i = tid; i < BW; i += NT) s_bitmap[i] = 0u;
      for (int i = tid; i < CB; i += NT) s_chist[i] = 0;
    }
    __syncthreads();
    const int n = m.start[R];
    if ((m.status & QF_QSTAT_BAD_LEN) || n == 0) {
      if (tid == 0) {
        p.out_lens[q] = 0;
        if (p.out_status) p.out_status[q] = m.status;
      }
      continue;
    }
    const int32_t* gid = p.doc_ids + (size_t)q * R * K;
    const raw_t* gsc = reinterpret_cast<const raw_t*>(p.scores) + (size_t)q * R * K;
    const int32_t* grk = (is_rrf && p.ranks) ? p.ranks + (size_t)q * R * K : nullptr;
    {   // pull the NEXT query of this CTA into L2 while this one is being fused
      const long long qn = (long long)q + gridDim.x;
      if (qn < p.Q) {
        const char* nid = reinterpret_cast<const char*>(p.doc_ids + (size_t)qn * R * K);
        const char* nsc = reinterpret_cast<const char*>(reinterpret_cast<const raw_t*>(p.scores) + (size_t)qn * R * K);
        const int id_lines = (R * K * 4 + 127) / 128;
        const int sc_lines = need_scores ? (R * K * (int)sizeof(raw_t) + 127) / 128 : 0;
        for (int l = tid; l < id_lines + sc_lines; l += NT) {
          const char* a = l < id_lines ? nid + (size_t)l * 128 : nsc + (size_t)(l - id_lines) * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
    }
    // ------------------------------------------------------------ ranx min-max: per-list min and range
    if (minmax) {
      for (int r = warp; r < R; r += NW) {
        const raw_t* gs = gsc + (size_t)r * K;
        const int len = m.len[r];
        if (zscore) {          // sequential float64 sums in list order: one lane per list (include/qppfusion.h)
          if (lane == 0 && len > 0) {
            double c = 0.0;
            for (int i = 0; i < len; ++i) c = __dadd_rn(c, raw_to_f64(__ldg(gs + i)));
            const double mean = __ddiv_rn(c, (double)len);
            double ssd = 0.0;
            for (int i = 0; i < len; ++i) {
              const double d = __dsub_rn(raw_to_f64(__ldg(gs + i)), mean);
              ssd = __dadd_rn(ssd, __dmul_rn(d, d));
            }
            const double sd = __dsqrt_rn(__ddiv_rn(ssd, (double)len));
            m.mn[r] = mean;
            m.den[r] = sd > 1e-9 ? sd : 1e-9;
          }
          continue;
        }
        double mn = INFINITY, mx = -INFINITY;
        for (int i = lane; i < len; i += 32) {
          const double v = raw_to_f64(__ldg(gs + i));
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) {
          m.mn[r] = mn;
          const double d = __dsub_rn(mx, mn);
          m.den[r] = d > 1e-9 ? d : 1e-9;     // ranx: max(max - min, 1e-9)
        }
      }
      __syncthreads();
    }

    // ------------------------------------------------------------ A1: join -- doc ids only, no ordering constraint
    // Every posting of every list finds / claims its doc's slot and records it in s_slot[r*K + i].  Nothing here
    // depends on the order in which postings arrive (the first-seen posting and the additions are settled in A2), so
    // there is no barrier between lists, the four postings of a thread are resolved side by side (straight-line
    // code, four independent dependency chains), and the postings whose home group holds four foreign keys or whose
    // claim was lost are only queued: they are resolved once per query on densely packed lanes.
    const bool vec = (K % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(gid) | reinterpret_cast<uintptr_t>(gsc) |
                       reinterpret_cast<uintptr_t>(grk)) & 15) == 0;
    auto load_ids = [&](int (&id)[4], int r, int base, int len) {
      const int32_t* gi = gid + (size_t)r * K;
      if (vec) {
        const int i0 = base + tid * 4;
        if (i0 < len) {                         // K % 4 == 0 and len <= K: the 4-wide load stays inside the row
          const int4 v = __ldg(reinterpret_cast<const int4*>(gi + i0));
          id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = base + j * NT + tid;
          if (i < len) id[j] = __ldg(gi + i);
        }
      }
    };
    // One probe step: classify the group with one select chain -- doc found (no atomic) / first empty key, claimed
    // with one CAS (a CAS on a key that meanwhile holds this very doc returns it = found) / four foreign keys or a
    // claim lost to another doc (-> not resolved: queued).  Keys only ever change from EMPTY to a doc id, and a doc
    // always sits before the first empty key of its probe order (home group, then groups at the doc's own stride;
    // views can only be stale towards "too empty", and then the CAS fails), so concurrent postings of the same doc
    // from different lists end up in the same slot.
    bool defer = false;
    int ins = 0;
    // unresolved postings -> CTA queue (one shared-memory atomic per warp and chunk)
    auto push_failed = [&](const bool (&fail)[4], u32 rowoff, int i0) {
      unsigned bal[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bal[j] = __ballot_sync(0xffffffffu, fail[j]);
      if (bal[0] | bal[1] | bal[2] | bal[3]) {
        const int c0 = __popc(bal[0]), c1 = __popc(bal[1]), c2 = __popc(bal[2]), c3 = __popc(bal[3]);
        u32 qb = 0;
        if (lane == 0) qb = atoms_add(qn_sa, (u32)(c0 + c1 + c2 + c3));
        qb = __shfl_sync(0xffffffffu, qb, 0);
        const unsigned lt = (1u << lane) - 1u;
        const u32 pre[4] = {qb, qb + (u32)c0, qb + (u32)(c0 + c1), qb + (u32)(c0 + c1 + c2)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const u32 qp = pre[j] + (u32)__popc(bal[j] & lt);
          if (fail[j] && qp < (u32)QCAP) s_queue[qp] = (u16)(rowoff + (u32)(vec ? i0 + j : i0 + j * NT));
        }
      }
    };
    auto store_slots = [&](const bool (&valid)[4], const u32 (&sl)[4], u32 rowoff, int i0) {
      if (vec) {                                           // 4 consecutive u16 entries, 8-byte aligned
        if (valid[0])
          *reinterpret_cast<uint2*>(s_slot + rowoff + (u32)i0) = make_uint2(sl[0] | (sl[1] << 16), sl[2] | (sl[3] << 16));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (valid[j]) s_slot[rowoff + (u32)(i0 + j * NT)] = (u16)sl[j];
      }
    };
    // The FIRST non-empty list meets an empty table and (barring a doc id that occurs twice in it) only new docs:
    // no probing at all -- a per-group fill counter (kept in the group's first meta word, cleared afterwards) hands
    // out the key positions 0..3, the rest of the group's postings are queued.  A doc id occurring twice would sit
    // in its group twice: checked group by group after the list's barrier.
    auto first4 = [&](const int (&id)[4], int r, int base, int len) {
      const u32 rowoff = (u32)r * (u32)K;
      const int i0 = vec ? base + tid * 4 : base + tid;
      bool valid[4], fail[4];
      u32 sl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        valid[j] = (vec ? i0 + j : i0 + j * NT) < len;
        defer |= valid[j] && (u32)id[j] == EMPTY;
        const u32 g = __umulhi((u32)id[j] * 2654435761u, (u32)G);
        const u32 c = atoms_add_if(valid[j], meta_sa + 16u * g, 1u, 4u);
        const bool ok = c < 4u;
        const u32 cand = 4u * g + c;
        if (ok) s_keys[cand] = (u32)id[j];
        ins += ok ? 1 : 0;
        fail[j] = valid[j] && !ok;
        sl[j] = ok ? cand : 0xFFFFu;
      }
      store_slots(valid, sl, rowoff, i0);
      push_failed(fail, rowoff, i0);
    };
    auto join4 = [&](const int (&id)[4], int r, int base, int len) {
      const u32 rowoff = (u32)r * (u32)K;
      const int i0 = vec ? base + tid * 4 : base + tid;
      bool valid[4], fail[4];
      u32 g[4], sl[4];
      uint4 kk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        valid[j] = (vec ? i0 + j : i0 + j * NT) < len;
        defer |= valid[j] && (u32)id[j] == EMPTY;          // doc id -1 is the empty-key sentinel: general kernel
        g[j] = __umulhi((u32)id[j] * 2654435761u, (u32)G);
        kk[j] = lds128_volatile(keys_sa + 16u * g[j]);     // (a dead lane reads some group: harmless)
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pos = probe4(kk[j], (u32)id[j]);
        const bool claim = valid[j] && pos >= 4;
        const u32 cand = 4u * g[j] + (u32)(pos & 3);
        const u32 o = atoms_cas_if(claim, keys_sa + 4u * cand, EMPTY, (u32)id[j], (u32)id[j]);
        const bool ok = pos >= 0 && (o == EMPTY || o == (u32)id[j]);
        ins += (claim && o == EMPTY) ? 1 : 0;
        fail[j] = valid[j] && !ok;
        sl[j] = ok ? cand : 0xFFFFu;
      }
      store_slots(valid, sl, rowoff, i0);
      push_failed(fail, rowoff, i0);
    };
    {
      const int r0 = m.first;
      int r = r0, base = 0, len_r = m.len[r];             // n > 0: there is a non-empty list
      int cur[4] = {0, 0, 0, 0}, nxt[4] = {0, 0, 0, 0};
      load_ids(cur, r, base, len_r);
      while (r < R) {
        int r2 = r, b2 = base + CH, len2 = len_r;
        if (b2 >= len_r) {
          b2 = 0;
          r2 = m.next[r];
          len2 = r2 < R ? m.len[r2] : 0;
        }
        if (r2 < R) load_ids(nxt, r2, b2, len2);
        if (r == r0) {
          first4(cur, r, base, len_r);
          if (r2 != r0) {                                  // first list complete: keys visible, counters cleared
            __syncthreads();
            for (int g = tid; g < G; g += NT) {
              const uint4 k = reinterpret_cast<const uint4*>(s_keys)[g];
              defer |= (k.x != EMPTY && (k.x == k.y || k.x == k.z || k.x == k.w)) ||
                       (k.y != EMPTY && (k.y == k.z || k.y == k.w)) || (k.z != EMPTY && k.z == k.w);
              s_meta[4 * g] = 0u;
            }
          }
        } else {
          join4(cur, r, base, len_r);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
        r = r2;
        base = b2;
        len_r = len2;
      }
    }

    // ------------------------------------------------------------ A2 prologue: first chunk of scores in flight
    struct SC {
      raw_t sc[4];
      int rk[4];
    };
    auto load_sc = [&](SC& c, int r, int base, int len) {
      const raw_t* gs = gsc + (size_t)r * K;
      if (vec) {
        const int i0 = base + tid * 4;
        if (i0 < len) {
          if (need_scores) {
            if constexpr (sizeof(raw_t) == 4) {
              const int4 s = __ldg(reinterpret_cast<const int4*>(gs + i0));
              c.sc[0] = *reinterpret_cast<const raw_t*>(&s.x);
              c.sc[1] = *reinterpret_cast<const raw_t*>(&s.y);
              c.sc[2] = *reinterpret_cast<const raw_t*>(&s.z);
              c.sc[3] = *reinterpret_cast<const raw_t*>(&s.w);
            } else {
              const double2 s0 = __ldg(reinterpret_cast<const double2*>(gs + i0));
              const double2 s1 = __ldg(reinterpret_cast<const double2*>(gs + i0) + 1);
              c.sc[0] = (raw_t)s0.x; c.sc[1] = (raw_t)s0.y; c.sc[2] = (raw_t)s1.x; c.sc[3] = (raw_t)s1.y;
            }
          }
          if (grk) {
            const int4 k4 = __ldg(reinterpret_cast<const int4*>(grk + (size_t)r * K + i0));
            c.rk[0] = k4.x; c.rk[1] = k4.y; c.rk[2] = k4.z; c.rk[3] = k4.w;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = base + j * NT + tid;
          if (i < len) {
            if (need_scores) c.sc[j] = __ldg(gs + i);
            if (grk) c.rk[j] = __ldg(grk + (size_t)r * K + i);
          }
        }
      }
    };
    const int r_first = m.first;
    SC cur_sc, nxt_sc;
    load_sc(cur_sc, r_first, 0, m.len[r_first]);

    // ------------------------------------------------------------ A1 drain: the queued postings, 32 per warp round
    __syncthreads();                                           // queue complete
    {
      const int qn = m.qn;
      if (qn > QCAP) {
        defer = true;                                          // (never at sane table loads) general kernel
      } else {
        for (int b = warp * 32; b < qn; b += NT) {
          const bool mine = b + lane < qn;
          u32 off = 0, id = 0, gq = 0, step = 1;
          int slot = -1;
          if (mine) {
            off = s_queue[b + lane];
            id = (u32)__ldg(gid + off);
            const u32 h = id * 2654435761u;
            gq = __umulhi(h, (u32)G);
            step = ((h >> 4) & 0xFEu) | 1u;                    // odd stride < 256 <= G: no primary clustering
          }
          uint4 k4 = lds128_volatile(keys_sa + 16u * gq);
          bool inserted = false, active = mine;
          int probes = 0;
          while (__any_sync(0xffffffffu, active)) {            // probe loop uniform over the warp
            if (active) {
              const int pos = probe4(k4, id);
              const u32 cand = 4u * gq + (u32)(pos & 3);
              const u32 o = atoms_cas_if(pos >= 4, keys_sa + 4u * cand, EMPTY, id, id);
              if (pos >= 0 && (o == EMPTY || o == id)) {
                inserted = (pos >= 4) && (o == EMPTY);
                slot = (int)cand;
                active = false;
              } else if (pos < 0) {                            // four foreign keys: next group of this doc's sequence
                if (++probes > MAXPROBE) {
                  defer = true;
                  active = false;
                }
                gq += step;
                gq = gq >= (u32)G ? gq - (u32)G : gq;
              }                                                // else: lost the key to another doc, same group again
              k4 = lds128_volatile(keys_sa + 16u * gq);
            }
          }
          if (slot >= 0) {
            s_slot[off] = (u16)slot;
            ins += inserted ? 1 : 0;
          }
        }
      }
    }
    {
      const int wi = warp_sum_int(ins);
      if (lane == 0 && wi) atomicAdd(&m.ins_total, wi);
    }
    const int stop = __syncthreads_or(defer ? 1 : 0);
    const int total_ins = m.ins_total;
    if (stop || total_ins > UCAP) {
      if (tid == 0) p.out_lens[q] = FUSE_DEFERRED;
      continue;
    }

    // ------------------------------------------------------------ A2: accumulate, list by list (one barrier per list)
    // Posting i of list r adds fl(w_r * s_i) to its doc's float64 accumulator with a plain read-modify-write: doc ids
    // are distinct inside a list and lists are separated by barriers, so the additions happen in exactly the
    // reference's order ((w0*s0 + w1*s1) + ...), no FMA contraction.  The posting that finds the slot's list mask
    // empty is the doc's FIRST-SEEN posting; its sequence number goes into the upper half of the meta word (tie-break
    // key).  No atomics: a doc id that occurs twice inside one list makes two postings set the same mask bit, so the
    // mask popcounts no longer add up to the number of postings -- checked in phase C, query deferred.
    bool bad_order = false;
    auto acc4 = [&](const SC& c, int r, int base, int len) {
      const int s0 = m.start[r];
      const double w = m.w[r];
      const double lmn = minmax ? m.mn[r] : 0.0, lden = minmax ? m.den[r] : 1.0;
      const u32 rowoff = (u32)r * (u32)K;
      int idx[4];
      bool valid[4];
      u32 slot[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        idx[j] = vec ? base + tid * 4 + j : base + j * NT + tid;
        valid[j] = idx[j] < len;
      }
      if (vec) {
        if (valid[0]) {                                        // 4 consecutive u16 slots, 8-byte aligned
          const uint2 s2 = *reinterpret_cast<const uint2*>(s_slot + rowoff + (u32)idx[0]);
          slot[0] = s2.x & 0xFFFFu; slot[1] = s2.x >> 16; slot[2] = s2.y & 0xFFFFu; slot[3] = s2.y >> 16;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (valid[j]) slot[j] = s_slot[rowoff + (u32)idx[j]];
      }
      if (check_sorted) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (valid[j] && idx[j] > 0) {
            const raw_t prev = (vec && j > 0) ? c.sc[j > 0 ? j - 1 : 0] : __ldg(gsc + (size_t)r * K + idx[j] - 1);
            if (prev < c.sc[j]) bad_order = true;
          }
      }
      if (grk) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (valid[j] && idx[j] > 0) {
            const int prev = (vec && j > 0) ? c.rk[j > 0 ? j - 1 : 0] : __ldg(grk + (size_t)r * K + idx[j] - 1);
            if (prev > c.rk[j]) bad_order = true;
          }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (valid[j]) {
          double v;
          bool counts = true;
          if (is_rrf) {
            const double rank = grk ? (double)c.rk[j] : (double)(idx[j] + 1);
            v = __ddiv_rn(w, __dadd_rn(p.rrf_k, rank));
          } else {
            double s = raw_to_f64(c.sc[j]);
            if (minmax) s = __ddiv_rn(__dsub_rn(s, lmn), lden);
            if (skip_zero && s == 0.0) counts = false;
            v = __dmul_rn(w, s);                             // w == 1.0 when unweighted: exact
          }
          u32 bits = 1u << r;
          if (skip_zero && counts) bits |= 1u << (8 + r);
          const u32 mt = s_meta[slot[j]];
          const bool first = (mt & 0xFFFFu) == 0u;           // no list has touched this doc yet
          if (first) {
            const u32 seq = (u32)(s0 + idx[j]);
            bits |= seq << 16;
            if (want_pos) atomicOr(&s_bitmap[seq >> 5], 1u << (seq & 31u));   // first-seen bitmap (out_pos)
          }
          s_meta[slot[j]] = mt | bits;
          const double prev = first ? 0.0 : s_acc[slot[j]];    // the reference starts from 0.0 (src/fusion.py:275)
          s_acc[slot[j]] = __dadd_rn(prev, v);
        }
      }
    };
    {
      int r = r_first, base = 0, len_r = m.len[r];
      while (r < R) {
        int r2 = r, b2 = base + CH, len2 = len_r;
        if (b2 >= len_r) {
          b2 = 0;
          r2 = m.next[r];
          len2 = r2 < R ? m.len[r2] : 0;
        }
        if (r2 < R) load_sc(nxt_sc, r2, b2, len2);
        acc4(cur_sc, r, base, len_r);
        if (r2 != r) __syncthreads();                         // list boundary
        cur_sc = nxt_sc;
        r = r2;
        base = b2;
        len_r = len2;
      }
    }
    if (bad_order) atomicOr(&m.status, QF_QSTAT_UNSORTED);
    for (int i = tid; i < NB / 2; i += NT) s_cnt32[i] = 0u;    // the bucket counters share memory with s_slot

    // ------------------------------------------------------------ B: score range and histogram from a sample
    // The union size is the number of claimed keys.  Final score of a slot = accumulator (x popcount for CombMNZ).
    // The accumulator started from +0.0 and x + (-x) rounds to +0.0, so it is never -0.0 (Python treats
    // -0.0 == 0.0 as a tie): nothing to canonicalise.
    // The bucket map only has to be MONOTONE in the score, so its range and its equalisation histogram come from a
    // sample (each thread's first occupied slot); scores outside the sampled range are clamped into the end buckets
    // and ordered there by the exact comparison like everything else.
    auto cntbits = [&](u32 mt) -> int { return __popc(skip_zero ? ((mt >> 8) & 0xFFu) : (mt & 0xFFFFu)); };
    auto final_score = [&](int slot, u32 mt) -> double {
      const double a = s_acc[slot];
      return is_mnz ? __dmul_rn(a, (double)cntbits(mt)) : a;
    };
    const int n_u = total_ins;
    double fmn = INFINITY, fmx = -INFINITY, samp = 0.0;
    bool has;
    {                                                       // sample: this thread's first occupied slot of two
      static_assert(H >= 2 * NT, "two sample probes per thread");
      const bool h0 = s_keys[tid] != EMPTY, h1 = s_keys[tid + NT] != EMPTY;
      const int slot = h0 ? tid : tid + NT;
      has = h0 | h1;
      if (has) samp = final_score(slot, s_meta[slot]);
    }
    auto block_range = [&](double lo, double hi) {          // -> fmn / fmx over the block (one barrier)
      lo = warp_min(lo);
      hi = warp_max(hi);
      if (lane == 0) {
        m.red_mn[warp] = lo;
        m.red_mx[warp] = hi;
      }
      __syncthreads();
      fmn = INFINITY;
      fmx = -INFINITY;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        fmn = m.red_mn[w] < fmn ? m.red_mn[w] : fmn;
        fmx = m.red_mx[w] > fmx ? m.red_mx[w] : fmx;
      }
    };
    block_range(has ? samp : INFINITY, has ? samp : -INFINITY);
    if (!unranked && !(fmx > fmn)) {
      // every sampled score equal: find out whether ALL scores are equal (then only first-seen order matters)
      double lo = INFINITY, hi = -INFINITY;
#pragma unroll 1
      for (int k = 0; k < SPT; ++k) {
        const int slot = tid + k * NT;
        if (slot < H && s_keys[slot] != EMPTY) {
          const double a = final_score(slot, s_meta[slot]);
          lo = a < lo ? a : lo;
          hi = a > hi ? a : hi;
        }
      }
      __syncthreads();                                       // red_mn / red_mx are being re-used
      block_range(lo, hi);
    }
    const bool score_mode = !unranked && (fmx > fmn);
    double cscale = 0.0;
    if (score_mode) {
      cscale = (double)CB / __dsub_rn(fmx, fmn);
      if (!(cscale > 0.0) || isinf(cscale)) cscale = 0.0;
    }
    auto coarse = [&](double fs, float& frac) -> int {
      float tf = (float)__dmul_rn(__dsub_rn(fmx, fs), cscale);           // monotone in fs
      tf = fminf(fmaxf(tf, 0.0f), (float)CB);                            // scores outside the sampled range
      int c = (int)tf;
      c = c > CB - 1 ? CB - 1 : c;
      frac = tf - (float)c;
      return c;
    };
    if (score_mode && has) {
      float fr;
      atomicAdd(&s_chist[coarse(samp, fr)], 1);
    }
    __syncthreads();
    // ------------------------------------------------------------ C: equalisation map, first-seen prefix
    if (warp == 0 && score_mode) {
      const int h0 = s_chist[2 * lane] + 1, h1 = s_chist[2 * lane + 1] + 1;
      const int tot = warp_sum_int(h0 + h1);
      const int w0 = (h0 * NB) / tot, w1 = (h1 * NB) / tot;
      int incl = w0 + w1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int excl = incl - (w0 + w1);
      s_ccum[2 * lane] = (float)excl;
      s_cw[2 * lane] = (float)w0;
      s_ccum[2 * lane + 1] = (float)(excl + w0);
      s_cw[2 * lane + 1] = (float)w1;
    }
    if (warp == 1 && want_pos) {
      int carry = 0;
      for (int b0 = 0; b0 < BW; b0 += 32) {
        const int wd = (b0 + lane < BW) ? __popc(s_bitmap[b0 + lane]) : 0;
        int incl = wd;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (b0 + lane < BW) s_bpre[b0 + lane] = (u16)(carry + incl - wd);
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
    const u32 tie_div = (u32)(want_pos ? n_u : n);       // range of the tie-break key in the meta word
    auto bucket = [&](double fs, u32 tie) -> int {
      int b;
      if (score_mode) {
        float frac;
        const int c = coarse(fs, frac);
        b = (int)fmaf(frac, s_cw[c], s_ccum[c]);          // monotone: cum[c] + frac*w[c] <= cum[c+1]
      } else {
        b = (int)((tie * (u32)NB) / tie_div);             // first-seen order only
      }
      return b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b);
    };
    // bucket counts; the atomicAdd's old value is the doc's offset inside its bucket
    u16 offs[SPT];
    int pc = 0;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int slot = tid + k * NT;
      offs[k] = 0xFFFFu;
      if (slot < H && s_keys[slot] != EMPTY) {
        const u32 mt = s_meta[slot];
        pc += __popc(mt & (skip_zero ? 0xFFu : 0xFFFFu));
        u32 tie = mt >> 16;
        if (want_pos) {      // first-seen sequence number -> first-seen index (same order, dense)
          tie = (u32)s_bpre[tie >> 5] + (u32)__popc(s_bitmap[tie >> 5] & ((1u << (tie & 31u)) - 1u));
          s_meta[slot] = (tie << 16) | (mt & 0xFFFFu);
        }
        double fs = s_acc[slot];
        if (is_mnz) {
          fs = __dmul_rn(fs, (double)cntbits(mt));
          s_acc[slot] = fs;
        }
        const int b = bucket(fs, tie);
        const int sh = (b & 1) * 16;
        const u32 oldc = atomicAdd(&s_cnt32[b >> 1], 1u << sh);
        s_bcache[slot] = (u16)b;
        offs[k] = (u16)((oldc >> sh) & 0xFFFFu);
      }
    }
    {
      const int wp = warp_sum_int(pc);
      if (lane == 0) atomicAdd(&m.pc, wp);
    }
    __syncthreads();
    // exclusive scan of the NB 16-bit counters (CPT consecutive counters per thread), in place
    {
      int tsum = 0, tmax = 0;
      uint4 cv[CPT / 8];
#pragma unroll
      for (int v = 0; v < CPT / 8; ++v) {
        cv[v] = reinterpret_cast<const uint4*>(s_cnt32)[tid * (CPT / 8) + v];
        u32 wds[4] = {cv[v].x, cv[v].y, cv[v].z, cv[v].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = (int)(wds[i] & 0xFFFFu), c1 = (int)(wds[i] >> 16);
          tmax = c0 > tmax ? c0 : tmax;
          tmax = c1 > tmax ? c1 : tmax;
          wds[i] = (u32)tsum | ((u32)(tsum + c0) << 16);       // exclusive prefixes inside the thread
          tsum += c0 + c1;
        }
        cv[v] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
      }
      int incl = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) m.warp_tot[warp] = incl;
      const int lng = __syncthreads_or((tmax > LONGB || m.pc != n) ? 1 : 0);
      if (lng) {                                // long run of (nearly) tied scores, or a doc id twice inside a list
        if (tid == 0) p.out_lens[q] = FUSE_DEFERRED;
        continue;
      }
      int wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += m.warp_tot[w];
      const u32 base2 = (u32)(wbase + incl - tsum) * 0x00010001u;     // added to both 16-bit halves (no carry: < 65536)
#pragma unroll
      for (int v = 0; v < CPT / 8; ++v) {
        cv[v].x += base2; cv[v].y += base2; cv[v].z += base2; cv[v].w += base2;
        reinterpret_cast<uint4*>(s_cnt32)[tid * (CPT / 8) + v] = cv[v];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int slot = tid + k * NT;
      if (offs[k] != 0xFFFFu) s_order[s_base16[s_bcache[slot]] + offs[k]] = (u16)slot;
    }
    __syncthreads();
    // ------------------------------------------------------------ D: exact rank inside the bucket, output
    {
      int32_t* oid = p.out_ids + (size_t)q * out_stride;
      double* osc = p.out_scores + (size_t)q * out_stride;
      asm volatile("" : "+l"(oid), "+l"(osc));           // keep the row pointers in registers (no per-element rebuild)
      for (int j = tid; j < n_u; j += NT) {
        const int slot = s_order[j];
        const double fs = s_acc[slot];
        const u32 mt = s_meta[slot];
        const u32 tie = mt >> 16;
        const int b = s_bcache[slot];
        const int lo = s_base16[b];
        const int hi = (b + 1 < NB) ? (int)s_base16[b + 1] : n_u;
        int rank = lo;
        if (score_mode) {
#pragma unroll 1
          for (int k = lo; k < hi; ++k) {
            const int s2 = s_order[k];
            const double f2 = s_acc[s2];
            const u32 t2 = s_meta[s2] >> 16;                 // loaded unconditionally: no branch inside the loop
            rank += (int)((f2 > fs) | ((f2 == fs) & (t2 < tie)));
          }
        } else {
          for (int k = lo; k < hi; ++k) rank += ((s_meta[s_order[k]] >> 16) < tie) ? 1 : 0;
        }
        oid[rank] = (int32_t)s_keys[slot];
        osc[rank] = fs;
        if (p.out_counts) p.out_counts[(size_t)q * out_stride + rank] = (u16)cntbits(mt);
        if (want_pos) p.out_pos[(size_t)q * out_stride + rank] = (u16)tie;
      }
      if (tid == 0) {
        p.out_lens[q] = n_u;
        if (p.out_status) p.out_status[q] = m.status;
      }
    }
  }
}

template <int G, int UCAP, int NB, int NMAX, int NT, int MINB, int ST, int MODE>
int launch_stream(const FuseParams& p, cudaStream_t stream) {
  using L = SLayout<G, UCAP, NB, NMAX, NT>;
  auto kern = fuse_stream_kernel<G, UCAP, NB, NMAX, NT, MINB, ST, MODE>;
  static bool configured_dev[64] = {};   // per instantiation; attribute set is idempotent
  bool& configured = configured_dev[qf::device_index()];   // kernel attributes are per device
  static int per_sm_dev[64] = {};
  int& per_sm = per_sm_dev[qf::device_index()];
  if (!configured) {
    QF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    QF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    QF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, L::total));
    if (per_sm < 1) per_sm = 1;
    configured = true;
  }
  long long grid = (long long)qf::sm_count() * per_sm;
  if (grid > p.Q) grid = p.Q;
  kern<<<(unsigned)grid, NT, L::total, stream>>>(p);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

template <int ST, int MODE>
int dispatch_stream(const FuseParams& p, cudaStream_t stream) {
  const long long n_max = (long long)p.R * p.K;
  if (n_max <= 1024) return launch_stream<384, 1024, 1024, 1024, 128, 5, ST, MODE>(p, stream);
  // Two table sizes.  Default: 3072 slots, 3 CTAs x 256 threads per SM -- fastest while a query's union stays below
  // ~2300 docs (table <= 75 % full), degrading quickly above.  QF_FLAG_LARGE_UNION: 5120 slots, 2 CTAs x 384 threads
  // -- ~10 % slower on small unions, 1.7-2.4x faster from ~2700 docs up, and nothing is deferred for size up to 4352.
  // QF_STREAM_VARIANT=0/1 (debug/tuning hook, read once) overrides the flag.
  static const int forced = [] {
    const char* e = getenv("QF_STREAM_VARIANT");
    return e ? atoi(e) : -1;
  }();
  const bool large = forced >= 0 ? forced == 1 : (p.flags & QF_FLAG_LARGE_UNION) != 0;
  if (large) return launch_stream<1280, 4352, 3072, 5120, 384, 2, ST, MODE>(p, stream);
  return launch_stream<768, 2816, 4096, 5120, 256, 3, ST, MODE>(p, stream);
}

template <int ST>
int dispatch_mode(const FuseParams& p, cudaStream_t stream) {
  const bool plain = p.method != QF_METHOD_RRF && p.norm == QF_NORM_NONE &&
                     !(p.flags & (QF_FLAG_MNZ_SKIP_ZERO | QF_FLAG_REQUIRE_SORTED));
  return plain ? dispatch_stream<ST, 0>(p, stream) : dispatch_stream<ST, 1>(p, stream);
}

}  // namespace

namespace qf {

bool fuse_stream_supported(const FuseParams& p) {
  if (p.flags & (FUSE_FLAG_DEBUG_FULLSORT | FUSE_FLAG_DEBUG_GENERAL)) return false;
  if (p.R > SMAX_R) return false;
  if ((p.flags & QF_FLAG_MNZ_SKIP_ZERO) && p.R > 8) return false;
  if ((long long)p.R * p.K > 5120) return false;
  return true;
}

int fuse_stream_launch(const FuseParams& p, int score_type, cudaStream_t stream) {
  switch (score_type) {
    case QF_SCORE_F64: return dispatch_mode<QF_SCORE_F64>(p, stream);
    case QF_SCORE_F32: return dispatch_mode<QF_SCORE_F32>(p, stream);
    default: return dispatch_mode<QF_SCORE_MICRO>(p, stream);
  }
}

}  // namespace qf
