// qf_fuse: rank fusion of R run lists per query, one CTA per query, everything resident in shared memory.
//
// Replaces ranx.fuse (reference src/fusion.py:182,201,226) and the weighted_* Python dict loops
// (src/fusion.py:266-278, 309-327, 361-377).  Bit-exact float64 semantics are described in
// include/qppfusion.h; the phases are
//   P0  stage doc ids + raw scores of the query's R lists (coalesced global -> shared)
//   P0b optional per-list min/max (ranx min-max norm) and sortedness validation
//   P1  doc-id join: open-addressing hash in shared memory built WITHOUT atomics
//       (write-if-empty, barrier, verify; losers probe on) -> slot_of[seq]
//   P3  accumulation list by list (= the reference's ranker order); inside one list doc ids are
//       distinct, so plain read-modify-write is race free; duplicates are detected by a
//       write/verify tag and routed to an exact sequential path
//   P4  final score (MNZ multiply), min/max of the order-preserving integer image of the scores
//   P2  first-seen flags + block scan -> union list in first-seen order
//   P5  ranking: stable LSD radix sort (8-bit digits, warp match ranking, no atomics) on the top
//       32 bits of a range-normalised descending key, then an exact fix-up of equal-prefix runs by
//       full float64 comparison (long runs fall back to a full 64-bit sort)
#include "qf_fuse_shared.cuh"
#include "qf_radix.cuh"

#include <stdlib.h>

namespace {

using u16 = unsigned short;
using u32 = unsigned int;
using u64 = unsigned long long;

constexpr u16 EMPTY16 = 0xFFFFu;
constexpr int MAX_R = qf::FUSE_MAX_R;
#ifndef QF_FUSE_NT
#define QF_FUSE_NT 512
#endif
constexpr u32 QF_FLAG_DEBUG_FULLSORT = qf::FUSE_FLAG_DEBUG_FULLSORT;
constexpr int FIXLIST = 1024;                    // fix-up mode: deferred queries collected per scan window

using qf::FuseParams;
using qf::ScoreT;
using qf::raw_to_f64;

// Shared-memory carve-up.  Region X (scores/table/slot_of/tag) is dead once P2 is done and is reused
// by the sort buffers.
template <int CAP, int HBITS, int NT, int ST>
struct Layout {
  static constexpr int H = 1 << HBITS;
  static constexpr int NW = NT / 32;
  using raw_t = typename ScoreT<ST>::type;
  static constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
  static constexpr size_t off_ids = 0;
  static constexpr size_t off_x = align16(off_ids + sizeof(int) * CAP);
  // region X, phase 1 view
  static constexpr size_t off_sc = off_x;
  static constexpr size_t off_table = align16(off_sc + sizeof(raw_t) * CAP);
  static constexpr size_t off_slotof = align16(off_table + sizeof(u16) * H);
  static constexpr size_t off_tag = align16(off_slotof + sizeof(u16) * CAP);
  static constexpr size_t end_x1 = align16(off_tag + sizeof(u16) * H);
  // region X, sort view
  static constexpr size_t off_keyA = off_x;
  static constexpr size_t off_keyB = align16(off_keyA + sizeof(u32) * CAP);
  static constexpr size_t off_idxA = align16(off_keyB + sizeof(u32) * CAP);
  static constexpr size_t off_idxB = align16(off_idxA + sizeof(u16) * CAP);
  static constexpr size_t end_x2 = align16(off_idxB + sizeof(u16) * CAP);
  static constexpr size_t end_x = end_x1 > end_x2 ? end_x1 : end_x2;
  static constexpr size_t off_first = end_x;
  static constexpr size_t off_cnt = align16(off_first + sizeof(u16) * H);
  static constexpr size_t off_acc = align16(off_cnt + sizeof(u16) * H);
  // the hash-join work lists (P1) live in the accumulator region, which is only initialised after P1
  static constexpr size_t off_wlA = off_acc;
  static constexpr size_t off_wlB = align16(off_wlA + sizeof(u32) * CAP);
  static_assert(2 * sizeof(u32) * CAP + 32 <= sizeof(double) * H, "work lists must fit in the accumulator region");
  static constexpr size_t off_uslot = align16(off_acc + sizeof(double) * H);
  static constexpr size_t off_whist = align16(off_uslot + sizeof(u16) * CAP);
  static constexpr size_t off_misc = align16(off_whist + sizeof(u16) * NW * 256);
  static constexpr size_t off_fix = off_misc + 2048;           // fix-up mode work list
  static constexpr size_t total = off_fix + sizeof(int) * FIXLIST;
};

struct Misc {
  int start[MAX_R + 1];
  double mn[MAX_R];
  double den[MAX_R];
  double w[MAX_R];
  int flag_dup;
  int flag_status;
  int flag_long;
  int flag_single;
  int wl_cnt;
  int fix_n;
  int vec_ok;
  int n_u;
  int warp_tot[32];
  u64 red_min[32];
  u64 red_max[32];
};
static_assert(sizeof(Misc) <= 2048, "Misc too large");

__device__ __forceinline__ int list_of(const int* start, int R, int seq) {
  int r = 0;
  while (r + 1 < R && seq >= start[r + 1]) ++r;
  return r;
}

template <int CAP, int HBITS, int NT, int ST>
__global__ void __launch_bounds__(NT) fuse_kernel(const FuseParams p) {
  using L = Layout<CAP, HBITS, NT, ST>;
  using raw_t = typename L::raw_t;
  constexpr int H = L::H;
  constexpr int NW = L::NW;
  constexpr int PPT = CAP / NT;          // postings per thread
  constexpr int MAXROUNDS = (CAP / 32 + NW - 1) / NW;
  static_assert(CAP % NT == 0 && CAP % 32 == 0, "CAP must be a multiple of NT");
  static_assert(CAP < 65535 && H <= 65536, "16-bit indices");

  extern __shared__ __align__(16) unsigned char smem[];
  int* s_ids = reinterpret_cast<int*>(smem + L::off_ids);
  raw_t* s_sc = reinterpret_cast<raw_t*>(smem + L::off_sc);
  u16* s_table = reinterpret_cast<u16*>(smem + L::off_table);
  u16* s_slotof = reinterpret_cast<u16*>(smem + L::off_slotof);
  u16* s_tag = reinterpret_cast<u16*>(smem + L::off_tag);
  u32* s_keyA = reinterpret_cast<u32*>(smem + L::off_keyA);
  u32* s_keyB = reinterpret_cast<u32*>(smem + L::off_keyB);
  u16* s_idxA = reinterpret_cast<u16*>(smem + L::off_idxA);
  u16* s_idxB = reinterpret_cast<u16*>(smem + L::off_idxB);
  u16* s_first = reinterpret_cast<u16*>(smem + L::off_first);
  u16* s_cnt = reinterpret_cast<u16*>(smem + L::off_cnt);
  double* s_acc = reinterpret_cast<double*>(smem + L::off_acc);
  u16* s_uslot = reinterpret_cast<u16*>(smem + L::off_uslot);
  u16* s_whist = reinterpret_cast<u16*>(smem + L::off_whist);
  u32* s_wlA = reinterpret_cast<u32*>(smem + L::off_wlA);
  u32* s_wlB = reinterpret_cast<u32*>(smem + L::off_wlB);
  Misc& m = *reinterpret_cast<Misc*>(smem + L::off_misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, K = p.K;
  const size_t out_stride = (size_t)R * K;

  // Normal mode: grid-stride over all queries.  Fix-up mode (p.fixup): this CTA owns a contiguous span of queries,
  // scans their out_lens in windows of FIXLIST and fuses only the ones the streaming kernel deferred.
  int* s_fix = reinterpret_cast<int*>(smem + L::off_fix);
  const int span = p.fixup ? (p.Q + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  int fix_win = 0, fix_k = 0, fix_n = 0;
  for (int it = 0;; ++it) {
    int q;
    if (!p.fixup) {
      q = blockIdx.x + it * gridDim.x;
      if (q >= p.Q) break;
    } else {
      while (fix_k >= fix_n && fix_win < span) {
        __syncthreads();
        if (tid == 0) m.fix_n = 0;
        __syncthreads();
        const int wend = min(span, fix_win + FIXLIST);
        for (int j = fix_win + tid; j < wend; j += NT) {
          const long long qq = (long long)blockIdx.x * span + j;
          if (qq < p.Q && p.out_lens[qq] == qf::FUSE_DEFERRED) s_fix[atomicAdd(&m.fix_n, 1)] = (int)qq;
        }
        __syncthreads();
        fix_n = m.fix_n;
        fix_k = 0;
        fix_win = wend;
      }
      if (fix_k >= fix_n) break;
      q = s_fix[fix_k++];
    }
    // ---------------------------------------------------------------- P0: stage the query's lists
    if (tid == 0) {
      int acc = 0, bad = 0, vec = (K % 4 == 0) ? 1 : 0;
      for (int r = 0; r < R; ++r) {
        m.start[r] = acc;
        const int len = p.lens[(size_t)q * R + r];
        if (len < 0 || len > K) bad = 1;
        if (len & 3) vec = 0;
        acc += bad ? 0 : len;
        m.w[r] = p.weights ? p.weights[(size_t)q * R + r] : 1.0;
      }
      m.start[R] = acc;
      m.flag_dup = 0;
      m.flag_long = 0;
      m.wl_cnt = 0;
      m.flag_status = bad ? QF_QSTAT_BAD_LEN : 0;
      m.vec_ok = vec;
    }
    {
      u32* t32 = reinterpret_cast<u32*>(s_table);
      for (int i = tid; i < H / 2; i += NT) t32[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    const int n = m.start[R];
    if (m.flag_status & QF_QSTAT_BAD_LEN) {
      if (tid == 0) {
        p.out_lens[q] = 0;
        if (p.out_status) p.out_status[q] = m.flag_status;
      }
      __syncthreads();
      continue;
    }
    {
      const int32_t* gid = p.doc_ids + (size_t)q * R * K;
      const raw_t* gsc = reinterpret_cast<const raw_t*>(p.scores) + (size_t)q * R * K;
      const bool vec = m.vec_ok && ((reinterpret_cast<uintptr_t>(gid) | reinterpret_cast<uintptr_t>(gsc)) & 15) == 0;
      if (vec) {
        // every list length is a multiple of 4: 16-byte loads, all of a thread's loads in flight together
        for (int c4 = tid * 4; c4 < n; c4 += NT * 4) {
          const int r = list_of(m.start, R, c4);
          const int i = c4 - m.start[r];
          *reinterpret_cast<int4*>(s_ids + c4) = __ldg(reinterpret_cast<const int4*>(gid + (size_t)r * K + i));
          if (sizeof(raw_t) == 4) {
            *reinterpret_cast<int4*>(s_sc + c4) = __ldg(reinterpret_cast<const int4*>(gsc + (size_t)r * K + i));
          } else {
            const int4* src = reinterpret_cast<const int4*>(gsc + (size_t)r * K + i);
            reinterpret_cast<int4*>(s_sc + c4)[0] = __ldg(src);
            reinterpret_cast<int4*>(s_sc + c4)[1] = __ldg(src + 1);
          }
        }
      } else {
        for (int r = 0; r < R; ++r) {
          const int s0 = m.start[r], len = m.start[r + 1] - s0;
          for (int i = tid; i < len; i += NT) {
            s_ids[s0 + i] = __ldg(gid + (size_t)r * K + i);
            s_sc[s0 + i] = __ldg(gsc + (size_t)r * K + i);
          }
        }
      }
      // pull the NEXT query of this CTA into L2 while this one is being fused (one 128-byte line per thread)
      const long long qn = p.fixup ? (long long)p.Q : (long long)q + gridDim.x;
      if (qn < p.Q) {
        const char* nid = reinterpret_cast<const char*>(p.doc_ids + (size_t)qn * R * K);
        const char* nsc = reinterpret_cast<const char*>(reinterpret_cast<const raw_t*>(p.scores) + (size_t)qn * R * K);
        const int id_lines = (R * K * 4 + 127) / 128, sc_lines = (R * K * (int)sizeof(raw_t) + 127) / 128;
        for (int l = tid; l < id_lines + sc_lines; l += NT) {
          const char* a = l < id_lines ? nid + (size_t)l * 128 : nsc + (size_t)(l - id_lines) * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- P0b: per-list min/max, validation
    if (p.norm == QF_NORM_MINMAX) {
      for (int r = warp; r < R; r += NW) {
        const int s0 = m.start[r], s1 = m.start[r + 1];
        double mn = INFINITY, mx = -INFINITY;
        for (int i = s0 + lane; i < s1; i += 32) {
          const double v = raw_to_f64(s_sc[i]);
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
        mn = qf::warp_min(mn);
        mx = qf::warp_max(mx);
        if (lane == 0) {
          m.mn[r] = mn;
          const double d = __dsub_rn(mx, mn);
          m.den[r] = d > 1e-9 ? d : 1e-9;     // ranx: max(max - min, 1e-9)
        }
      }
    } else if (p.norm == QF_NORM_ZSCORE) {
      // sequential float64 sums in list order (one lane per list: the order of the additions is the definition)
      for (int r = warp; r < R; r += NW) {
        const int s0 = m.start[r], s1 = m.start[r + 1];
        if (lane == 0 && s1 > s0) {
          double c = 0.0;
          for (int i = s0; i < s1; ++i) c = __dadd_rn(c, raw_to_f64(s_sc[i]));
          const double mean = __ddiv_rn(c, (double)(s1 - s0));
          double ssd = 0.0;
          for (int i = s0; i < s1; ++i) {
            const double d = __dsub_rn(raw_to_f64(s_sc[i]), mean);
            ssd = __dadd_rn(ssd, __dmul_rn(d, d));
          }
          const double sd = __dsqrt_rn(__ddiv_rn(ssd, (double)(s1 - s0)));
          m.mn[r] = mean;
          m.den[r] = sd > 1e-9 ? sd : 1e-9;
        }
      }
    }
    {
      int bad = 0;
      if (p.flags & QF_FLAG_REQUIRE_SORTED) {
        for (int seq = tid; seq < n; seq += NT) {
          const int r = list_of(m.start, R, seq);
          if (seq > m.start[r] && raw_to_f64(s_sc[seq - 1]) < raw_to_f64(s_sc[seq])) bad = 1;
        }
      }
      if (p.method == QF_METHOD_RRF && p.ranks) {
        const int32_t* grk = p.ranks + (size_t)q * R * K;
        for (int seq = tid; seq < n; seq += NT) {
          const int r = list_of(m.start, R, seq);
          const int i = seq - m.start[r];
          if (i > 0 && __ldg(grk + (size_t)r * K + i - 1) > __ldg(grk + (size_t)r * K + i)) bad = 1;
        }
      }
      if (bad) m.flag_status = QF_QSTAT_UNSORTED;   // benign race: all writers store the same value
    }

    // ---------------------------------------------------------------- P1: hash join (no atomics on the table)
    // Round 0 visits every posting; postings that lose their slot to a DIFFERENT doc id go to a work list
    // (seq | next probe position << 16) and only those are revisited in later rounds.  All postings of one
    // doc id probe the same sequence in lock-step, so they settle in the same slot.
    {
      u32* wl_out = s_wlA;
      u32* wl_in = s_wlB;
      u16 hpos[PPT];
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int seq = tid + k * NT;
        if (seq < n) {
          hpos[k] = (u16)(((u32)s_ids[seq] * 2654435761u) >> (32 - HBITS));
          if (s_table[hpos[k]] == EMPTY16) s_table[hpos[k]] = (u16)seq;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int seq = tid + k * NT;
        if (seq < n) {
          const u16 owner = s_table[hpos[k]];
          if (s_ids[owner] == s_ids[seq]) {
            s_slotof[seq] = hpos[k];
          } else {
            const int pos = atomicAdd(&m.wl_cnt, 1);
            wl_out[pos] = (u32)seq | ((u32)((hpos[k] + 1) & (H - 1)) << 16);
          }
        }
      }
      __syncthreads();
      while (true) {
        const int cnt = m.wl_cnt;
        if (cnt == 0) break;
        { u32* t = wl_in; wl_in = wl_out; wl_out = t; }
        __syncthreads();                       // everyone has read wl_cnt
        if (tid == 0) m.wl_cnt = 0;
        for (int e = tid; e < cnt; e += NT) {
          const u32 ent = wl_in[e];
          const u32 h = ent >> 16;
          if (s_table[h] == EMPTY16) s_table[h] = (u16)(ent & 0xFFFFu);
        }
        __syncthreads();
        for (int e = tid; e < cnt; e += NT) {
          const u32 ent = wl_in[e];
          const u32 h = ent >> 16, seq = ent & 0xFFFFu;
          const u16 owner = s_table[h];
          if (s_ids[owner] == s_ids[seq]) {
            s_slotof[seq] = (u16)h;
          } else {
            const int pos = atomicAdd(&m.wl_cnt, 1);
            wl_out[pos] = seq | (((h + 1) & (H - 1)) << 16);
          }
        }
        __syncthreads();
      }
    }
    // per-slot accumulators (their region held the work lists until now)
    {
      u32* f32 = reinterpret_cast<u32*>(s_first);
      u32* c32 = reinterpret_cast<u32*>(s_cnt);
      for (int i = tid; i < H / 2; i += NT) {
        f32[i] = 0xFFFFFFFFu;
        c32[i] = 0u;
      }
      double2* a2 = reinterpret_cast<double2*>(s_acc);
      for (int i = tid; i < H / 2; i += NT) a2[i] = make_double2(0.0, 0.0);
    }
    __syncthreads();

    // ---------------------------------------------------------------- P3: ordered accumulation
    const bool is_rrf = (p.method == QF_METHOD_RRF);
    const bool weighted = (p.weights != nullptr);
    const bool skip_zero = (p.flags & QF_FLAG_MNZ_SKIP_ZERO) != 0;
    const int32_t* grk = p.ranks ? p.ranks + (size_t)q * R * K : nullptr;

    auto contribution = [&](int seq, int r, bool& counts) -> double {
      counts = true;
      if (is_rrf) {
        const int i = seq - m.start[r];
        const double rank = grk ? (double)__ldg(grk + (size_t)r * K + i) : (double)(i + 1);
        return __ddiv_rn(m.w[r], __dadd_rn(p.rrf_k, rank));
      }
      double s = raw_to_f64(s_sc[seq]);
      if (p.norm != QF_NORM_NONE) s = __ddiv_rn(__dsub_rn(s, m.mn[r]), m.den[r]);   // min-max: mn = min; z-score: mn = mean
      if (skip_zero && s == 0.0) counts = false;
      return weighted ? __dmul_rn(m.w[r], s) : s;
    };

    for (int r = 0; r < R; ++r) {
      const int s0 = m.start[r], s1 = m.start[r + 1];
      for (int seq = s0 + tid; seq < s1; seq += NT) s_tag[s_slotof[seq]] = (u16)seq;
      __syncthreads();
      for (int seq = s0 + tid; seq < s1; seq += NT) {
        const int slot = s_slotof[seq];
        if (s_tag[slot] != (u16)seq) {
          m.flag_dup = 1;
        } else {
          bool counts;
          const double v = contribution(seq, r, counts);
          if (s_first[slot] == EMPTY16) s_first[slot] = (u16)seq;
          s_acc[slot] = __dadd_rn(s_acc[slot], v);
          if (counts) s_cnt[slot] = (u16)(s_cnt[slot] + 1);
        }
      }
      __syncthreads();
    }
    if (m.flag_dup) {
      // exact sequential path: a doc id repeats inside one list (reference: every row accumulates
      // and counts, src/fusion.py:275,323-324)
      for (int i = tid; i < H; i += NT) {
        s_first[i] = EMPTY16;
        s_cnt[i] = 0;
        s_acc[i] = 0.0;
      }
      __syncthreads();
      if (tid == 0) {
        int r = 0;
        for (int seq = 0; seq < n; ++seq) {
          while (seq >= m.start[r + 1]) ++r;
          const int slot = s_slotof[seq];
          bool counts;
          const double v = contribution(seq, r, counts);
          if (s_first[slot] == EMPTY16) s_first[slot] = (u16)seq;
          s_acc[slot] = __dadd_rn(s_acc[slot], v);
          if (counts && s_cnt[slot] != 0xFFFFu) s_cnt[slot] = (u16)(s_cnt[slot] + 1);
        }
        m.flag_status |= QF_QSTAT_DUPLICATES | ((p.flags & QF_FLAG_REQUIRE_SORTED) ? QF_QSTAT_UNSORTED : 0);
      }
      __syncthreads();
    }

    // ---------------------------------------------------------------- P2: first-seen order
    {
      const int base_seq = tid * PPT;
      int cnt = 0;
      u32 firstmask = 0;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const int seq = base_seq + k;
        if (seq < n && s_first[s_slotof[seq]] == (u16)seq) {
          ++cnt;
          firstmask |= 1u << k;
        }
      }
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) m.warp_tot[warp] = incl;
      __syncthreads();
      int wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += m.warp_tot[w];
      int u = wbase + incl - cnt;
#pragma unroll
      for (int k = 0; k < PPT; ++k)
        if (firstmask & (1u << k)) s_uslot[u++] = s_slotof[base_seq + k];
      if (tid == NT - 1) m.n_u = u;
    }
    __syncthreads();   // region X (scores/table/slot_of/tag) is dead from here on
    const int n_u = m.n_u;

    // ---------------------------------------------------------------- P4: final scores, key range (over the union)
    u64 omin = ~0ull, omax = 0ull;
    {
      for (int u = tid; u < n_u; u += NT) {
        const int slot = s_uslot[u];
        double fs = s_acc[slot];
        if (p.method == QF_METHOD_MNZ) fs = __dmul_rn(fs, (double)s_cnt[slot]);
        fs = fs + 0.0;                       // -0.0 -> +0.0 (Python: -0.0 == 0.0 is a tie)
        s_acc[slot] = fs;
        const u64 o = qf::ord_f64(fs);
        omin = o < omin ? o : omin;
        omax = o > omax ? o : omax;
      }
      omin = qf::warp_min(omin);
      omax = qf::warp_max(omax);
      if (lane == 0) {
        m.red_min[warp] = omin;
        m.red_max[warp] = omax;
      }
      __syncthreads();
      omin = ~0ull;
      omax = 0ull;
      for (int w = 0; w < NW; ++w) {
        omin = m.red_min[w] < omin ? m.red_min[w] : omin;
        omax = m.red_max[w] > omax ? m.red_max[w] : omax;
      }
    }

    // ---------------------------------------------------------------- P5: ranking
    const u16* s_rank = nullptr;     // idx array holding the final order
    const int n_pad = (n_u + 31) & ~31;
    const bool want_rank = !(p.flags & QF_FLAG_UNRANKED_OUTPUT) && n_u > 1 && omax != omin;
    if (!want_rank) {
      for (int u = tid; u < n_u; u += NT) s_idxA[u] = (u16)u;
      __syncthreads();
      s_rank = s_idxA;
    } else {
      const int lz = __clzll((long long)(omax - omin));
      auto key64 = [&](int u) -> u64 { return (omax - qf::ord_f64(s_acc[s_uslot[u]])) << lz; };
      u16* idxIn = s_idxA;
      u16* idxOut = s_idxB;
      bool full_sort = (p.flags & QF_FLAG_DEBUG_FULLSORT) != 0;
      if (!full_sort) {
        for (int u = tid; u < n_pad; u += NT) {
          s_keyA[u] = u < n_u ? (u32)(key64(u) >> 32) : 0xFFFFFFFFu;
          s_idxA[u] = (u16)u;
        }
        __syncthreads();
        u32* keyIn = s_keyA;
        u32* keyOut = s_keyB;
#pragma unroll 1
        for (int pass = 1; pass < 4; ++pass) {       // top 24 bits of the key: 3 passes, the fix-up is exact
          const int shift = pass * 8;
          const u32* kin = keyIn;
          auto dig = [&](int e) -> u32 { return (kin[e] >> shift) & 255u; };
          if (qf::radix_pass<NT, MAXROUNDS, true>(dig, n_pad, keyIn, keyOut, idxIn, idxOut, s_whist, m.warp_tot, &m.flag_single)) {
            u32* tk = keyIn; keyIn = keyOut; keyOut = tk;
            u16* ti = idxIn; idxIn = idxOut; idxOut = ti;
          }
        }
        // exact fix-up of runs whose sorted 24-bit prefix is equal
        for (int j = tid; j < n_u; j += NT) {
          const u32 kj = keyIn[j] >> 8;
          const bool head = (j == 0 || (keyIn[j - 1] >> 8) != kj) && (j + 1 < n_u && (keyIn[j + 1] >> 8) == kj);
          if (head) {
            int e = j + 2;
            while (e < n_u && (keyIn[e] >> 8) == kj) ++e;
            if (e - j > 48) {
              m.flag_long = 1;
            } else {
              for (int a = j + 1; a < e; ++a) {       // stable insertion sort, score descending
                const u16 ia = idxIn[a];
                const double sa = s_acc[s_uslot[ia]];
                int b = a - 1;
                while (b >= j && s_acc[s_uslot[idxIn[b]]] < sa) {
                  idxIn[b + 1] = idxIn[b];
                  --b;
                }
                idxIn[b + 1] = ia;
              }
            }
          }
        }
        __syncthreads();
        full_sort = (m.flag_long != 0);
        if (full_sort) {            // restart from first-seen order
          idxIn = s_idxA;
          idxOut = s_idxB;
        }
      }
      if (full_sort) {
        __syncthreads();
        for (int u = tid; u < n_pad; u += NT) s_idxA[u] = (u16)u;
        __syncthreads();
#pragma unroll 1
        for (int pass = 0; pass < 8; ++pass) {
          const int shift = pass * 8;
          const u16* iin = idxIn;
          auto dig = [&](int e) -> u32 {
            const int u = iin[e];
            return u < n_u ? (u32)((key64(u) >> shift) & 255ull) : 255u;
          };
          // padding elements (u >= n_u) carry digit 255 in every pass and start at the tail, so the
          // stable passes keep them behind every real element
          if (qf::radix_pass<NT, MAXROUNDS, false>(dig, n_pad, nullptr, nullptr, idxIn, idxOut, s_whist, m.warp_tot, &m.flag_single)) {
            u16* ti = idxIn; idxIn = idxOut; idxOut = ti;
          }
        }
      }
      s_rank = idxIn;
    }

    // ---------------------------------------------------------------- output
    {
      int32_t* oid = p.out_ids + (size_t)q * out_stride;
      double* osc = p.out_scores + (size_t)q * out_stride;
      for (int j = tid; j < n_u; j += NT) {
        const int u = s_rank[j];
        const int slot = s_uslot[u];
        oid[j] = s_ids[s_first[slot]];
        osc[j] = s_acc[slot];
        if (p.out_counts) p.out_counts[(size_t)q * out_stride + j] = s_cnt[slot];
        if (p.out_pos) p.out_pos[(size_t)q * out_stride + j] = (u16)u;
      }
      if (tid == 0) {
        p.out_lens[q] = n_u;
        if (p.out_status) p.out_status[q] = m.flag_status;
      }
    }
    __syncthreads();
  }
}

template <int CAP, int HBITS, int NT, int ST>
int launch(const FuseParams& p, cudaStream_t stream) {
  using L = Layout<CAP, HBITS, NT, ST>;
  auto kern = fuse_kernel<CAP, HBITS, NT, ST>;
  static bool configured_dev[64] = {};   // per instantiation; attribute set is idempotent
  bool& configured = configured_dev[qf::device_index()];   // kernel attributes are per device
  if (!configured) {
    QF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    configured = true;
  }
  int per_sm = 1;
  QF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, L::total));
  if (per_sm < 1) per_sm = 1;
  const int sms = qf::sm_count();
  long long grid = (long long)sms * per_sm;
  if (grid > p.Q) grid = p.Q;
  kern<<<(unsigned)grid, NT, L::total, stream>>>(p);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

// 6 144 postings per query (e.g. 6 retrievers x TREC depth 1000) fit the 227 KB of shared memory only with 4-byte
// scores (micro-units / float32): the float64 instantiation stops at 5 120.
constexpr int CAP_WIDE = 6144;
static_assert(Layout<CAP_WIDE, 13, QF_FUSE_NT, QF_SCORE_MICRO>::total <= 232448, "wide general kernel must fit in shared memory");

template <int ST>
int dispatch_cap(const FuseParams& p, cudaStream_t stream) {
  const long long n_max = (long long)p.R * p.K;
  if (n_max <= 1024) return launch<1024, 11, 128, ST>(p, stream);
  if (n_max <= 5120) return launch<5120, 13, QF_FUSE_NT, ST>(p, stream);
  if constexpr (ST == QF_SCORE_F64) {
    qf::set_error("qf_fuse: R*K=%lld exceeds 5120 postings per query for float64 scores (6144 with QF_SCORE_MICRO / "
                  "QF_SCORE_F32)", n_max);
    return QF_E_UNSUPPORTED;
  } else {
    return launch<CAP_WIDE, 13, QF_FUSE_NT, ST>(p, stream);
  }
}

}  // namespace

extern "C" int qf_fuse_max_postings(void) { return CAP_WIDE; }

extern "C" int qf_fuse(const int32_t* doc_ids, const void* scores, int score_type, const int32_t* ranks,
                       const int32_t* lens, const double* weights, int Q, int R, int K, int method, int norm,
                       double rrf_k, uint32_t flags, int32_t* out_ids, double* out_scores, uint16_t* out_counts,
                       uint16_t* out_pos, int32_t* out_lens, int32_t* out_status, void* stream) {
  QF_REQUIRE(Q >= 0 && R >= 1 && K >= 1, QF_E_INVALID, "qf_fuse: bad sizes Q=%d R=%d K=%d", Q, R, K);
  QF_REQUIRE(R <= MAX_R, QF_E_UNSUPPORTED, "qf_fuse: R=%d exceeds %d", R, MAX_R);
  QF_REQUIRE((long long)R * K <= qf_fuse_max_postings(), QF_E_UNSUPPORTED,
             "qf_fuse: R*K=%lld exceeds the shared-memory envelope of %d postings per query", (long long)R * K,
             qf_fuse_max_postings());
  QF_REQUIRE(method == QF_METHOD_SUM || method == QF_METHOD_MNZ || method == QF_METHOD_RRF, QF_E_INVALID,
             "qf_fuse: unknown method %d", method);
  QF_REQUIRE(norm == QF_NORM_NONE || norm == QF_NORM_MINMAX || norm == QF_NORM_ZSCORE, QF_E_INVALID,
             "qf_fuse: unknown norm %d", norm);
  QF_REQUIRE(score_type >= QF_SCORE_F64 && score_type <= QF_SCORE_MICRO, QF_E_INVALID,
             "qf_fuse: unknown score_type %d", score_type);
  if (Q == 0) return QF_OK;
  QF_REQUIRE(doc_ids && scores && lens && out_ids && out_scores && out_lens, QF_E_INVALID,
             "qf_fuse: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  FuseParams p{doc_ids, scores, ranks, lens, weights, Q, R, K, method, norm, rrf_k, flags,
               out_ids, out_scores, out_counts, out_pos, out_lens, out_status, 0};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool gather = qf::fuse_gather_supported(p, score_type);
  if (gather || qf::fuse_stream_supported(p)) {
    // common case: gather kernel (or the r1h streaming kernel outside its envelope); the general kernel then
    // finishes whatever it deferred (usually nothing)
    qf::prof_mark(QF_PROF_FUSE_STREAM, 0, st);
    rc = gather ? qf::fuse_gather_launch(p, score_type, st) : qf::fuse_stream_launch(p, score_type, st);
    qf::prof_mark(QF_PROF_FUSE_STREAM, 1, st);
    if (rc != QF_OK) return rc;
    p.fixup = 1;
    static const bool no_fixup = getenv("QF_STREAM_NOFIXUP") != nullptr;   // debug hook: leave deferred queries marked
    if (no_fixup) return QF_OK;
  }
  qf::prof_mark(QF_PROF_FUSE_GENERAL, 0, st);
  switch (score_type) {
    case QF_SCORE_F64: rc = dispatch_cap<QF_SCORE_F64>(p, st); break;
    case QF_SCORE_F32: rc = dispatch_cap<QF_SCORE_F32>(p, st); break;
    default: rc = dispatch_cap<QF_SCORE_MICRO>(p, st); break;
  }
  qf::prof_mark(QF_PROF_FUSE_GENERAL, 1, st);
  return rc;
}
