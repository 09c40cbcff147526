// qf_fuse_gather: rank fusion, "deposit + gather" design (round 2).  Same contract and results as qf_fuse_stream.cu
// (bit-exact float64, include/qppfusion.h); built around SHARED-MEMORY WAVEFRONTS instead of instructions: the r1h
// kernel spent 4 random shared-memory accesses per posting on the ordered float64 read-modify-write (accumulator +
// list mask) and ranked through an index array, i.e. random reads in the rank loop.  Here
//
//   stage   the query's score (or rank) rows are copied global -> shared by the TMA unit (cp.async.bulk, one copy per
//           list, mbarrier completion): no instruction, no register, no LSU wavefront; lands while the join runs.
//   join    order-free, doc ids only (as r1h: 4-key groups, one 16-byte probe, predicated CAS claim, first list through
//           per-group fill counters, stragglers queued and drained on dense lanes).  A resolved posting (r, i) DEPOSITS
//           its list position: P[r][slot] = i (one random 2-byte store; the only per-posting random access besides the
//           probe).  No accumulators, no list barriers, no s_slot array.
//   gather  per SLOT, by the thread that owns the slot's key group (consecutive lanes = consecutive addresses: keys and
//           P rows are read conflict-free with 16 / 8-byte loads): for r = 0..R-1 in the reference's order, if the doc
//           is in list r gather its score S[r][P[r][slot]] and add fl(w_r * s) -- sequential float64 adds in a
//           register, no atomics, no barriers between lists.  The first list that holds the doc gives the first-seen
//           sequence number, the number of lists the MNZ count; a doc id occurring twice inside one list overwrote a
//           P entry, so the counts no longer sum to the number of postings -> query deferred (general kernel).
//           Final scores, counts and first-seen keys stay in REGISTERS (12 slots per thread).
//   rank    exact min/max from the registers, 64-bin sample histogram -> monotone equalised map score -> 4096 buckets,
//           bucket counts by shared-memory atomicAdd (old value = offset in the bucket), block scan, then every slot
//           scatters a RECORD (float64 score | first-seen, count, bucket | doc id; three arrays) to its bucket-ordered
//           position.  The records overlay the dead P / S arrays.  Rank loop and output run over the dense record
//           arrays: a doc's bucket mates are its neighbours, consecutive lanes read consecutive addresses; the loop
//           compares float64 keys only, the first-seen tie-break is evaluated only for docs that met an equal score.
//
// Shared memory per query (R <= 5, 3 072 slots): keys 12 KB + P 30 KB + S 20 KB + counters 8 KB ~ 72 KB -> 3 CTAs x 256
// threads per SM (as r1h).  Envelope: R <= 8, K <= 1024 (list positions are 10 bits), R*K <= 5120, 4-byte scores
// (float32 / micro-units); everything else goes to qf_fuse_stream.cu.  Deferred queries (out_lens = FUSE_DEFERRED) are
// finished by qf_fuse.cu's general kernel in the same stream.
#include "qf_fuse_shared.cuh"

#include <stdlib.h>

namespace {

using namespace qf;
using u16 = unsigned short;
using u32 = unsigned int;
using u64 = unsigned long long;

constexpr u32 EMPTY = 0xFFFFFFFFu;
constexpr u32 PNONE = 0xFFFFu;
constexpr int GR = 8;            // lists per query this kernel is compiled for at most (P rows)
constexpr int KMAX = 1024;       // list positions are stored in 10 bits (queue entries: r << 10 | i)
constexpr int CB = 64;           // coarse bins of the equalisation histogram
constexpr int MAXPROBE = 64;     // key groups visited before the query is deferred
constexpr int LONGB = 160;       // longest bucket ranked by pairwise comparison

__device__ __forceinline__ uint4 lds128_volatile(u32 sa) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
  return v;
}
__device__ __forceinline__ u32 atoms_cas_if(bool pred, u32 sa, u32 cmp, u32 val, u32 dflt) {
  u32 old = dflt;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p atom.shared.cas.b32 %0, [%1], %2, %3;\n\t}"
               : "+r"(old) : "r"(sa), "r"(cmp), "r"(val), "r"((u32)pred) : "memory");
  return old;
}
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
__device__ __forceinline__ void sts16(u32 sa, u32 v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa), "h"((u16)v) : "memory");
}
// ---- mbarrier + bulk copy (TMA unit, no tensor map: 1-D, 16-byte granularity) ----
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct GMisc {
  int start[GR + 1];
  int len[GR];
  int next[GR];     // next non-empty list after r (R if none)
  int first;        // first non-empty list
  double w[GR], mn[GR], den[GR];
  double red_mn[16], red_mx[16];
  int warp_tot[16];
  int status;
  int qn;           // queued (unresolved) postings of the join pass
  int ins_total;    // claimed keys = size of the union
  int pc;           // sum over the slots of the number of lists holding the doc (== postings unless an id repeats in a list)
  u64 bar;          // mbarrier of the staging copies
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

// G key groups (4 keys each), RM list rows of P, NMAX postings per query, NB ranking buckets, UCAP largest union
template <int G, int NT, int RM, int NMAX, int NB, int UCAP, int QCAP>
struct GLayout {
  static constexpr int H = 4 * G;
  static constexpr int BW = NMAX / 32;                          // first-seen bitmap words
  static constexpr int CPT = NB / NT;                           // bucket counters scanned per thread
  static constexpr int GPT = (G + NT - 1) / NT;                 // key groups owned per thread
  static constexpr size_t a16(size_t x) { return (x + 15) & ~size_t(15); }
  static constexpr size_t off_keys = 0;
  static constexpr size_t off_P = off_keys + 4 * H;             // u16 P[RM][H]: list position of the doc in list r
  static constexpr size_t off_S = off_P + 2 * (size_t)RM * H;   // 4-byte staged scores / ranks, rows of Kp
  static constexpr size_t off_cnt = off_S + 4 * (size_t)NMAX;   // u16 bucket counters -> bases
  static constexpr size_t off_fill = off_cnt;                   // join only: u32 fill counter per key group
  static constexpr size_t off_queue = off_fill + 4 * G;         // join only: u16 queue of unresolved postings
  static constexpr size_t cnt_bytes = (2 * NB > 4 * G + 2 * QCAP) ? 2 * NB : 4 * G + 2 * QCAP;
  static constexpr size_t off_coarse = a16(off_cnt + cnt_bytes);      // int hist[CB] + float2 map[CB]
  static constexpr size_t off_bitmap = a16(off_coarse + 12 * CB);     // first-seen bitmap + word prefix (out_pos)
  static constexpr size_t off_misc = a16(off_bitmap + 6 * BW);
  static constexpr size_t total = a16(off_misc + sizeof(GMisc));
  // records (after the gather pass; overlay P and S)
  static constexpr size_t off_rkey = off_P;                     // double[UCAP]
  static constexpr size_t off_rmeta = off_rkey + 8 * (size_t)UCAP;    // u32[UCAP]: tie | count << 13 | bucket << 17
  static constexpr size_t off_rid = off_rmeta + 4 * (size_t)UCAP;     // u32[UCAP]
  static_assert(off_rid + 4 * (size_t)UCAP <= off_cnt, "records must fit into the P + S arrays");
  static_assert((4 * H + 2 * RM * H) % 16 == 0 && cnt_bytes % 16 == 0, "16-byte memset");
  static_assert(NB % NT == 0 && CPT % 8 == 0 && NB <= 4096, "bucket scan: CPT counters per thread; 12-bit bucket ids");
  static_assert(UCAP < H && UCAP < 8192 && NMAX % 32 == 0 && NMAX <= 8192 && H < 65536, "layout (13-bit tie keys)");
  static_assert(NT / 32 <= 16 && NT / 32 >= 2 && RM <= GR, "warps / lists");
};

__device__ __forceinline__ int warp_sum_int(int v) { return __reduce_add_sync(0xffffffffu, v); }

// MODE 0: CombSUM / CombMNZ (weighted or not) on raw scores -- no norm, no zero-skipping, no order checks, no RRF
//         (the weighted_* path and the benchmark step); MODE 1: every option decided at run time.
// ST: QF_SCORE_F32 or QF_SCORE_MICRO (4-byte scores).
template <int G, int NT, int MINB, int RM, int NMAX, int NB, int UCAP, int QCAP, int ST, int MODE>
__global__ void __launch_bounds__(NT, MINB) fuse_gather_kernel(const FuseParams p) {
  using L = GLayout<G, NT, RM, NMAX, NB, UCAP, QCAP>;
  using raw_t = typename ScoreT<ST>::type;
  static_assert(sizeof(raw_t) == 4, "4-byte scores");
  constexpr int H = L::H, NW = NT / 32, BW = L::BW, CPT = L::CPT, GPT = L::GPT, SPT = 4 * GPT;

  extern __shared__ __align__(16) unsigned char smem[];
  u32* s_keys = reinterpret_cast<u32*>(smem + L::off_keys);
  u16* s_P = reinterpret_cast<u16*>(smem + L::off_P);
  u32* s_S = reinterpret_cast<u32*>(smem + L::off_S);
  u32* s_cnt32 = reinterpret_cast<u32*>(smem + L::off_cnt);
  u16* s_base16 = reinterpret_cast<u16*>(smem + L::off_cnt);
  u16* s_queue = reinterpret_cast<u16*>(smem + L::off_queue);
  int* s_chist = reinterpret_cast<int*>(smem + L::off_coarse);
  float2* s_cmap = reinterpret_cast<float2*>(smem + L::off_coarse + 4 * CB);
  u32* s_bitmap = reinterpret_cast<u32*>(smem + L::off_bitmap);
  u16* s_bpre = reinterpret_cast<u16*>(smem + L::off_bitmap + 4 * BW);
  double* s_rkey = reinterpret_cast<double*>(smem + L::off_rkey);
  u32* s_rmeta = reinterpret_cast<u32*>(smem + L::off_rmeta);
  u32* s_rid = reinterpret_cast<u32*>(smem + L::off_rid);
  GMisc& m = *reinterpret_cast<GMisc*>(smem + L::off_misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = p.R, K = p.K;
  const int Kp = (K + 3) & ~3;                         // row stride of the staged array
  const size_t out_stride = (size_t)R * K;
  const u32 keys_sa = (u32)__cvta_generic_to_shared(s_keys);
  const u32 P_sa = (u32)__cvta_generic_to_shared(s_P);
  const u32 S_sa = (u32)__cvta_generic_to_shared(s_S);
  const u32 fill_sa = (u32)__cvta_generic_to_shared(smem + L::off_fill);
  const u32 bar_sa = (u32)__cvta_generic_to_shared(&m.bar);
  const u32 qn_sa = (u32)__cvta_generic_to_shared(&m.qn);
  const bool is_rrf = MODE ? (p.method == QF_METHOD_RRF) : false;
  const bool is_mnz = (p.method == QF_METHOD_MNZ);
  const bool weighted = (p.weights != nullptr);
  const bool skip_zero = MODE ? (p.flags & QF_FLAG_MNZ_SKIP_ZERO) != 0 : false;
  const bool check_sorted = MODE ? (p.flags & QF_FLAG_REQUIRE_SORTED) != 0 : false;
  const bool unranked = (p.flags & QF_FLAG_UNRANKED_OUTPUT) != 0;
  const bool minmax = MODE ? ((p.norm == QF_NORM_MINMAX) && !is_rrf) : false;
  const bool has_rk = MODE ? (is_rrf && p.ranks != nullptr) : false;
  const bool stage = MODE ? (!is_rrf || has_rk || check_sorted) : true;   // anything to stage at all
  const bool want_pos = (p.out_pos != nullptr);
  // what is staged: the rank column for RRF with explicit ranks, the scores otherwise
  const u32* stage_base = has_rk ? reinterpret_cast<const u32*>(p.ranks) : reinterpret_cast<const u32*>(p.scores);
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.doc_ids) | reinterpret_cast<uintptr_t>(stage_base)) & 15) == 0;

  if (tid == 0) {
    mbar_init(bar_sa, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  u32 tma_parity = 0;

  // warp 0 keeps the NEXT query's list lengths and weights in registers (loaded while the current one is fused)
  int pre_len = 0;
  double pre_w = 1.0;
  if (warp == 0 && lane < R && (int)blockIdx.x < p.Q) {
    pre_len = __ldg(p.lens + (size_t)blockIdx.x * R + lane);
    if (weighted) pre_w = __ldg(p.weights + (size_t)blockIdx.x * R + lane);
  }

  for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
    // ------------------------------------------------------------ header + reset
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
      h_l = (h_badm || lane >= R) ? 0 : len;
      h_incl = h_l;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, h_incl, o);
        if (lane >= o) h_incl += t;
      }
      h_nz = __ballot_sync(0xffffffffu, lane < R && h_l > 0);
    }
    __syncthreads();                       // previous query's shared state is dead
    const u32* stage_q = stage_base + (size_t)q * R * K;
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
      // staging copies: one bulk copy per non-empty list, issued by the list's lane, completing on the mbarrier
      if (stage && vec && h_nz) {
        const u32 bytes = ((u32)h_l * 4u + 15u) & ~15u;        // K % 4 == 0 and len <= K: stays inside the row
        const u32 total = (u32)warp_sum_int((int)bytes);
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores before, TMA writes after
          mbar_expect_tx(bar_sa, total);
        }
        __syncwarp();
        if (bytes) bulk_g2s(S_sa + 4u * (u32)(lane * Kp), stage_q + (size_t)lane * K, bytes, bar_sa);
      }
    }
    {
      uint4* z = reinterpret_cast<uint4*>(smem + L::off_keys);           // keys and P: all ones (EMPTY / PNONE)
      constexpr int NZ = (int)((4 * H + 2 * (size_t)RM * H) / 16);
      for (int i = tid; i < NZ; i += NT) z[i] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
      uint4* c = reinterpret_cast<uint4*>(smem + L::off_fill);            // fill counters
      for (int i = tid; i < G / 4; i += NT) c[i] = make_uint4(0u, 0u, 0u, 0u);
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
    const bool tma_issued = stage && vec;            // n > 0: at least one copy is in flight
    const int32_t* gid = p.doc_ids + (size_t)q * R * K;
    {   // pull the NEXT query's doc ids into L2 while this one is being fused
      const long long qx = (long long)q + gridDim.x;
      if (qx < p.Q) {
        const char* nid = reinterpret_cast<const char*>(p.doc_ids + (size_t)qx * R * K);
        const char* nsc = reinterpret_cast<const char*>(stage_base + (size_t)qx * R * K);
        const int id_lines = (R * K * 4 + 127) / 128;
        const int sc_lines = stage ? id_lines : 0;
        for (int l = tid; l < id_lines + sc_lines; l += NT) {
          const char* a = l < id_lines ? nid + (size_t)l * 128 : nsc + (size_t)(l - id_lines) * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
    }
    if (stage && !vec) {                             // unaligned rows: plain copy, visible after the join's barriers
      for (int r = 0; r < R; ++r) {
        const int len = m.len[r];
        for (int i = tid; i < len; i += NT) s_S[r * Kp + i] = __ldg(stage_q + (size_t)r * K + i);
      }
    }

    // ------------------------------------------------------------ join: doc ids only, no ordering constraint
    // NT * 4 >= KMAX >= K: every list is ONE chunk of at most 4 postings per thread.
    // (dispatch_gather guarantees K <= 4 * NT)
    auto load_ids = [&](int (&id)[4], int r, int len) {
      const int32_t* gi = gid + (size_t)r * K;
      if (vec) {
        const int i0 = tid * 4;
        if (i0 < len) {
          const int4 v = __ldg(reinterpret_cast<const int4*>(gi + i0));
          id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = j * NT + tid;
          if (i < len) id[j] = __ldg(gi + i);
        }
      }
    };
    bool defer = false;
    int ins = 0;
    auto push_failed = [&](const bool (&fail)[4], u32 rtag, int i0) {
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
          if (fail[j] && qp < (u32)QCAP) s_queue[qp] = (u16)(rtag | (u32)(vec ? i0 + j : i0 + j * NT));
        }
      }
    };
    // the FIRST non-empty list meets an empty table and only new docs: no probing, a per-group fill counter hands out
    // the key positions 0..3, the rest of the group's postings are queued
    auto first4 = [&](const int (&id)[4], int r, int len) {
      const u32 prow = P_sa + 2u * (u32)(r * H);
      const int i0 = vec ? tid * 4 : tid;
      bool fail[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = vec ? i0 + j : i0 + j * NT;
        const bool valid = i < len;
        defer |= valid && (u32)id[j] == EMPTY;
        const u32 g = __umulhi((u32)id[j] * 2654435761u, (u32)G);
        const u32 c = atoms_add_if(valid, fill_sa + 4u * g, 1u, 4u);
        const bool ok = c < 4u;
        const u32 cand = 4u * g + c;
        if (ok) {
          s_keys[cand] = (u32)id[j];
          sts16(prow + 2u * cand, (u32)i);
        }
        ins += ok ? 1 : 0;
        fail[j] = valid && !ok;
      }
      push_failed(fail, (u32)r << 10, i0);
    };
    auto join4 = [&](const int (&id)[4], int r, int len) {
      const u32 prow = P_sa + 2u * (u32)(r * H);
      const int i0 = vec ? tid * 4 : tid;
      bool valid[4], fail[4];
      u32 g[4];
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
        if (valid[j] && ok) sts16(prow + 2u * cand, (u32)(vec ? i0 + j : i0 + j * NT));
      }
      push_failed(fail, (u32)r << 10, i0);
    };
    {
      // all lists' doc ids in flight at once (RM x 4 registers; one 16-byte load per list and thread)
      const int r0 = m.first;
      int idv[RM][4];
      int lenr[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        lenr[r] = r < R ? m.len[r] : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) idv[r][j] = 0;
        if (lenr[r] > 0) load_ids(idv[r], r, lenr[r]);
      }
#pragma unroll
      for (int r = 0; r < RM; ++r)
        if (r == r0) first4(idv[r], r, lenr[r]);
      __syncthreads();                                     // first list complete: keys visible
      for (int g = tid; g < G; g += NT) {                  // a doc id twice in the first list sits in its group twice
        const uint4 k = reinterpret_cast<const uint4*>(s_keys)[g];
        defer |= (k.x != EMPTY && (k.x == k.y || k.x == k.z || k.x == k.w)) ||
                 (k.y != EMPTY && (k.y == k.z || k.y == k.w)) || (k.z != EMPTY && k.z == k.w);
      }
#pragma unroll
      for (int r = 0; r < RM; ++r)
        if (r != r0 && lenr[r] > 0) join4(idv[r], r, lenr[r]);
    }

    // ------------------------------------------------------------ drain: the queued postings, 32 per warp round
    __syncthreads();                                           // queue complete
    {
      const int qn = m.qn;
      if (qn > QCAP) {
        defer = true;                                          // (never at sane table loads) general kernel
      } else {
        for (int b = warp * 32; b < qn; b += NT) {
          const bool mine = b + lane < qn;
          u32 e = 0, id = 0, gq = 0, step = 1;
          int slot = -1;
          if (mine) {
            e = s_queue[b + lane];
            id = (u32)__ldg(gid + (size_t)(e >> 10) * K + (e & 1023u));
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
            sts16(P_sa + 2u * ((e >> 10) * (u32)H + (u32)slot), e & 1023u);
            ins += inserted ? 1 : 0;
          }
        }
      }
    }
    {
      const int wi = warp_sum_int(ins);
      if (lane == 0 && wi) atomicAdd(&m.ins_total, wi);
    }
    const int stop = __syncthreads_or(defer ? 1 : 0);          // table and P complete
    const int n_u = m.ins_total;
    if (tma_issued) {                                          // (every thread observes the phase: visibility of the
      mbar_wait(bar_sa, tma_parity);                           //  TMA writes, and the parity bit stays in step)
      tma_parity ^= 1u;
    }
    if (stop || n_u > UCAP) {
      if (tid == 0) p.out_lens[q] = FUSE_DEFERRED;
      continue;
    }
    // the fill counters / queue are dead: zero the bucket counters that share their memory
    {
      uint4* c = reinterpret_cast<uint4*>(smem + L::off_cnt);
      for (int i = tid; i < NB / 8; i += NT) c[i] = make_uint4(0u, 0u, 0u, 0u);
    }

    // ------------------------------------------------------------ per-list statistics on the staged rows (MODE 1)
    bool bad_order = false;
    if (MODE && (minmax || check_sorted || has_rk)) {
      for (int r = warp; r < R; r += NW) {
        const int len = m.len[r];
        const u32* row = s_S + r * Kp;
        double mn = INFINITY, mx = -INFINITY;
        for (int i = lane; i < len; i += 32) {
          const u32 raw = row[i];
          if (has_rk) {
            if (i > 0 && (int)row[i - 1] > (int)raw) bad_order = true;
          } else {
            const raw_t sv = *reinterpret_cast<const raw_t*>(&raw);
            if (check_sorted && i > 0) {
              const u32 pr = row[i - 1];
              if (*reinterpret_cast<const raw_t*>(&pr) < sv) bad_order = true;
            }
            if (minmax) {
              const double v = raw_to_f64(sv);
              mn = fmin(mn, v);
              mx = fmax(mx, v);
            }
          }
        }
        if (minmax) {
          mn = warp_min(mn);
          mx = warp_max(mx);
          if (lane == 0) {
            m.mn[r] = mn;
            const double d = __dsub_rn(mx, mn);
            m.den[r] = d > 1e-9 ? d : 1e-9;     // ranx: max(max - min, 1e-9)
          }
        }
      }
      if (bad_order) atomicOr(&m.status, QF_QSTAT_UNSORTED);
      if (minmax) __syncthreads();
    }

    // ------------------------------------------------------------ gather: per slot, lists in the reference's order
    double fsv[SPT];
    u32 mtv[SPT];                      // first-seen key | count << 13 ; 0xFFFFFFFF = empty slot
    double fmn = INFINITY, fmx = -INFINITY, samp = 0.0;
    bool has = false;
    {
      // An absent (slot, list) pair must cost as little as a present one saves: two thirds of the pairs are absent,
      // and every warp-level instruction below is issued for them too.  Per pair: halfword extract, compare, address,
      // predicated 4-byte gather, conversion, multiply, predicated add, predicated mask update.
      double wr[RM], lmn[RM], lden[RM];
      u32 st[RM], rowS[RM];
#pragma unroll
      for (int r = 0; r < RM; ++r) {
        wr[r] = r < R ? m.w[r] : 0.0;
        st[r] = r < R ? (u32)m.start[r] : 0u;
        rowS[r] = S_sa + 4u * (u32)(r * Kp);
        lmn[r] = (MODE && minmax && r < R) ? m.mn[r] : 0.0;
        lden[r] = (MODE && minmax && r < R) ? m.den[r] : 1.0;
      }
      int pc = 0;
#pragma unroll
      for (int gi = 0; gi < GPT; ++gi) {
        const int g = tid + gi * NT;
        uint4 k4 = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
        uint2 pr[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) pr[r] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        if (G % NT == 0 || g < G) {
          k4 = reinterpret_cast<const uint4*>(s_keys)[g];
#pragma unroll
          for (int r = 0; r < RM; ++r)
            if (r < R) pr[r] = *reinterpret_cast<const uint2*>(s_P + (size_t)r * H + 4 * g);
        }
        const u32 kj[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // (an empty slot has no P entries: nobody ever resolved a posting to it)
          double acc = 0.0;                                // the reference starts from 0.0 (src/fusion.py:275)
          u32 pm = 0u, pz = 0u;                            // lists holding the doc / with a non-zero score
          u32 ev[RM];
#pragma unroll
          for (int r = 0; r < RM; ++r) {
            ev[r] = PNONE;
            if (r < R) {
              const u32 e = __byte_perm((j < 2) ? pr[r].x : pr[r].y, 0u, (j & 1) ? 0x4432u : 0x4410u);
              ev[r] = e;
              const bool pres = e != PNONE;
              u32 raw = 0u;
              if (MODE == 0 || stage) {
                if (pres) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(rowS[r] + 4u * e));
              }
              double v;
              if (MODE && is_rrf) {
                const double rank = has_rk ? (double)(int)raw : (double)((int)e + 1);
                v = __ddiv_rn(wr[r], __dadd_rn(p.rrf_k, rank));
              } else {
                double sc = raw_to_f64(*reinterpret_cast<const raw_t*>(&raw));
                if (MODE && minmax) sc = __ddiv_rn(__dsub_rn(sc, lmn[r]), lden[r]);
                if (MODE && skip_zero && pres && sc != 0.0) pz |= 1u << r;
                v = __dmul_rn(wr[r], sc);                  // w == 1.0 when unweighted: exact
              }
              if (pres) {
                acc = __dadd_rn(acc, v);
                pm |= 1u << r;
              }
            }
          }
          u32 seq = 0u;                                    // first list that holds the doc: descending, last write wins
#pragma unroll
          for (int r = RM - 1; r >= 0; --r)
            if (r < R && ev[r] != PNONE) seq = st[r] + ev[r];
          const int cnt = __popc(pm);
          const int c_out = (MODE && skip_zero) ? __popc(pz) : cnt;
          const double fs = is_mnz ? __dmul_rn(acc, (double)c_out) : acc;
          // the accumulator started from +0.0 and x + (-x) rounds to +0.0, so it is never -0.0 (Python treats
          // -0.0 == 0.0 as a tie, and so do the comparisons below)
          const bool occ = kj[j] != EMPTY;
          fsv[gi * 4 + j] = fs;
          mtv[gi * 4 + j] = occ ? (seq | ((u32)c_out << 13)) : 0xFFFFFFFFu;
          pc += cnt;
          if (occ) {
            fmn = fmin(fmn, fs);
            fmx = fmax(fmx, fs);
            samp = fs;
            has = true;
            if (want_pos) atomicOr(&s_bitmap[seq >> 5], 1u << (seq & 31u));   // first-seen bitmap (out_pos)
          }
        }
      }
      const int wp = warp_sum_int(pc);
      if (lane == 0) atomicAdd(&m.pc, wp);
    }

    // ------------------------------------------------------------ score range (exact), sample histogram
    {
      const double lo = warp_min(fmn), hi = warp_max(fmx);
      if (lane == 0) {
        m.red_mn[warp] = lo;
        m.red_mx[warp] = hi;
      }
      __syncthreads();                                     // also: gather pass done, P / S dead; m.pc complete
      fmn = INFINITY;
      fmx = -INFINITY;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        fmn = m.red_mn[w] < fmn ? m.red_mn[w] : fmn;
        fmx = m.red_mx[w] > fmx ? m.red_mx[w] : fmx;
      }
    }
    if (m.pc != n) {                                       // a doc id twice inside one list (uniform over the CTA)
      if (tid == 0) p.out_lens[q] = FUSE_DEFERRED;
      continue;
    }
    const bool score_mode = !unranked && (fmx > fmn);
    double cscale = 0.0;
    if (score_mode) {
      cscale = (double)CB / __dsub_rn(fmx, fmn);
      if (!(cscale > 0.0) || isinf(cscale)) cscale = 0.0;
    }
    auto coarse = [&](double fs, float& frac) -> int {
      float tf = (float)__dmul_rn(__dsub_rn(fmx, fs), cscale);           // monotone in fs
      tf = fminf(fmaxf(tf, 0.0f), (float)CB);
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
    // ------------------------------------------------------------ equalisation map, first-seen prefix
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
      s_cmap[2 * lane] = make_float2((float)excl, (float)w0);
      s_cmap[2 * lane + 1] = make_float2((float)(excl + w0), (float)w1);
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
    const u32 tie_div = (u32)(want_pos ? n_u : n);       // range of the tie-break key
    // ------------------------------------------------------------ bucket counts; the atomicAdd's old value is the
    // doc's offset inside its bucket.  bo = bucket | offset << 12
    u32 bo[SPT];
    int omax = 0;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      bo[k] = 0u;
      if (mtv[k] != 0xFFFFFFFFu) {
        u32 tie = mtv[k] & 0x1FFFu;
        if (want_pos) {      // first-seen sequence number -> first-seen index (same order, dense)
          tie = (u32)s_bpre[tie >> 5] + (u32)__popc(s_bitmap[tie >> 5] & ((1u << (tie & 31u)) - 1u));
          mtv[k] = (mtv[k] & ~0x1FFFu) | tie;
        }
        int b;
        if (score_mode) {
          float frac;
          const int c = coarse(fsv[k], frac);
          const float2 cm = s_cmap[c];
          b = (int)fmaf(frac, cm.y, cm.x);                // monotone: cum[c] + frac*w[c] <= cum[c+1]
        } else {
          b = (int)((tie * (u32)NB) / tie_div);           // first-seen order only
        }
        b = b < 0 ? 0 : (b > NB - 1 ? NB - 1 : b);
        const int sh = (b & 1) * 16;
        const u32 oldc = atomicAdd(&s_cnt32[b >> 1], 1u << sh);
        const int off = (int)((oldc >> sh) & 0xFFFFu);
        omax = off > omax ? off : omax;
        bo[k] = (u32)b | ((u32)off << 12);
      }
    }
    const int lng = __syncthreads_or(omax >= LONGB ? 1 : 0);
    if (lng) {                                  // long run of (nearly) tied scores
      if (tid == 0) p.out_lens[q] = FUSE_DEFERRED;
      continue;
    }
    // exclusive scan of the NB 16-bit counters (CPT consecutive counters per thread), in place
    {
      int tsum = 0;
      uint4 cv[CPT / 8];
#pragma unroll
      for (int v = 0; v < CPT / 8; ++v) {
        cv[v] = reinterpret_cast<const uint4*>(s_cnt32)[tid * (CPT / 8) + v];
        u32 wds[4] = {cv[v].x, cv[v].y, cv[v].z, cv[v].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = (int)(wds[i] & 0xFFFFu), c1 = (int)(wds[i] >> 16);
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
      __syncthreads();
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
    // ------------------------------------------------------------ scatter the records into bucket order
#pragma unroll
    for (int gi = 0; gi < GPT; ++gi) {
      const int g = tid + gi * NT;
      uint4 k4 = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
      if (G % NT == 0 || g < G) k4 = reinterpret_cast<const uint4*>(s_keys)[g];
      const u32 kj[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = gi * 4 + j;
        if (mtv[k] != 0xFFFFFFFFu) {
          const u32 b = bo[k] & 0xFFFu;
          const int pos = (int)s_base16[b] + (int)(bo[k] >> 12);
          s_rkey[pos] = fsv[k];
          s_rmeta[pos] = (mtv[k] & 0x1FFFFu) | (b << 17);
          s_rid[pos] = kj[j];
        }
      }
    }
    __syncthreads();
    // ------------------------------------------------------------ exact rank inside the bucket, output
    {
      int32_t* oid = p.out_ids + (size_t)q * out_stride;
      double* osc = p.out_scores + (size_t)q * out_stride;
      asm volatile("" : "+l"(oid), "+l"(osc));           // keep the row pointers in registers
      for (int j = tid; j < n_u; j += NT) {
        const double fs = s_rkey[j];
        const u32 mt = s_rmeta[j];
        const u32 tie = mt & 0x1FFFu;
        const int b = (int)(mt >> 17);
        const int lo = s_base16[b];
        const int hi = (b + 1 < NB) ? (int)s_base16[b + 1] : n_u;
        int rank = lo;
        if (score_mode) {
          int eq = 0;
#pragma unroll 1
          for (int k = lo; k < hi; ++k) {
            const double f2 = s_rkey[k];
            rank += (int)(f2 > fs);
            eq += (int)(f2 == fs);
          }
          if (eq > 1) {                                    // equal scores met: first-seen order decides
#pragma unroll 1
            for (int k = lo; k < hi; ++k) rank += (int)((s_rkey[k] == fs) & ((s_rmeta[k] & 0x1FFFu) < tie));
          }
        } else {
          for (int k = lo; k < hi; ++k) rank += ((s_rmeta[k] & 0x1FFFu) < tie) ? 1 : 0;
        }
        oid[rank] = (int32_t)s_rid[j];
        osc[rank] = fs;
        if (p.out_counts) p.out_counts[(size_t)q * out_stride + rank] = (u16)((mt >> 13) & 0xFu);
        if (want_pos) p.out_pos[(size_t)q * out_stride + rank] = (u16)tie;
      }
      if (tid == 0) {
        p.out_lens[q] = n_u;
        if (p.out_status) p.out_status[q] = m.status;
      }
    }
  }
}

template <int G, int NT, int MINB, int RM, int NMAX, int NB, int UCAP, int QCAP, int ST, int MODE>
int launch_gather(const FuseParams& p, cudaStream_t stream) {
  using L = GLayout<G, NT, RM, NMAX, NB, UCAP, QCAP>;
  auto kern = fuse_gather_kernel<G, NT, MINB, RM, NMAX, NB, UCAP, QCAP, ST, MODE>;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[qf::device_index()];
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
int dispatch_gather(const FuseParams& p, cudaStream_t stream) {
  const long long n_max = (long long)p.R * p.K;
  if (n_max <= 1024 && p.K <= 512) return launch_gather<384, 128, 5, 8, 1024, 1024, 1024, 512, ST, MODE>(p, stream);
  // QF_GATHER_VARIANT=0/1 (debug/tuning hook, read once) overrides the choice below.
  static const int forced = [] {
    const char* e = getenv("QF_GATHER_VARIANT");
    return e ? atoi(e) : -1;
  }();
  const bool large = forced >= 0 ? forced == 1 : ((p.flags & QF_FLAG_LARGE_UNION) != 0 || p.R > 5);
  // default: 3 072 slots, R <= 5, 3 CTAs x 256 threads per SM; large: 4 096 slots, R <= 8, 2 CTAs x 384 threads
  if (large || p.R > 5) return launch_gather<1024, 384, 2, 8, 5120, 3072, 3840, 2048, ST, MODE>(p, stream);
  return launch_gather<768, 256, 3, 5, 5120, 4096, 2816, 2560, ST, MODE>(p, stream);
}

template <int ST>
int dispatch_gmode(const FuseParams& p, cudaStream_t stream) {
  const bool plain = p.method != QF_METHOD_RRF && p.norm == QF_NORM_NONE &&
                     !(p.flags & (QF_FLAG_MNZ_SKIP_ZERO | QF_FLAG_REQUIRE_SORTED));
  return plain ? dispatch_gather<ST, 0>(p, stream) : dispatch_gather<ST, 1>(p, stream);
}

}  // namespace

namespace qf {

bool fuse_gather_supported(const FuseParams& p, int score_type) {
  static const bool on = getenv("QF_FUSE_GATHER") != nullptr;        // A/B hook: gather kernel wherever it applies
  if (!on && !(p.flags & FUSE_FLAG_GATHER)) return false;
  if (p.flags & (FUSE_FLAG_DEBUG_FULLSORT | FUSE_FLAG_DEBUG_GENERAL)) return false;
  if (score_type != QF_SCORE_F32 && score_type != QF_SCORE_MICRO) return false;
  if (p.R > GR || p.K > KMAX) return false;
  if ((long long)p.R * p.K > 5120) return false;
  if (p.norm != QF_NORM_NONE && p.norm != QF_NORM_MINMAX) return false;
  // RRF on the file's rank column stages the ranks; validating the score order as well would need both arrays
  if (p.method == QF_METHOD_RRF && p.ranks != nullptr && (p.flags & QF_FLAG_REQUIRE_SORTED)) return false;
  return true;
}

int fuse_gather_launch(const FuseParams& p, int score_type, cudaStream_t stream) {
  return score_type == QF_SCORE_F32 ? dispatch_gmode<QF_SCORE_F32>(p, stream) : dispatch_gmode<QF_SCORE_MICRO>(p, stream);
}

}  // namespace qf
