// qf_dense_topk: exact inner-product top-k of a bf16 query batch against a bf16 corpus matrix
// (reference: BGERetriever._search_faiss_batch -> faiss.IndexFlatIP.search, src/retrievers/bge.py:179-194).
//
// B200 design: the GEMM  S[Q,N] = Qm[Q,D] . C[N,D]^T  runs on the 5th-gen tensor cores and the score
// matrix never reaches HBM:
//   * a persistent CTA keeps ONE 128-query tile resident in TENSOR MEMORY as the A operand
//     (128 lanes x D/2 columns, D <= 768 -> 384 of the 512 TMEM columns); the remaining 128 columns are
//     two 64-column fp32 accumulator stages
//   * the corpus streams through shared memory exactly once per query tile: TMA (cp.async.bulk.tensor,
//     SWIZZLE_128B) fills a ring of 24 x 8 KB boxes (64 docs x 64 k), an mbarrier pipeline feeds
//     tcgen05.mma.cta_group::1.kind::f16 (A from TMEM, B from smem descriptors, M=128 N=64 K=16)
//   * four epilogue warps read each finished accumulator stage with tcgen05.ld (TMEM lane == query row,
//     so every thread owns one query's threshold in a register), and either
//        FILTER  : append (score, doc) pairs with score >= tau[q] to the query's candidate list, or
//        STORE   : write the raw scores (threshold-estimation sample / exact fallback)
//   * tau[q] comes from an exact order statistic of a strided corpus sample; a per-query select kernel
//     then ranks the few thousand candidates (score desc, id asc) and emits the top k.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "qf_common.cuh"

#include <stdlib.h>
#include "qf_radix.cuh"

namespace {

using qf::u16;
using qf::u32;
using qf::u64;

constexpr int BM = 128;                 // queries per tile (TMEM lanes)
constexpr int BN = 64;                  // docs per accumulator stage
constexpr int BK = 64;                  // k elements per TMA box (128 B rows, SWIZZLE_128B)
constexpr int MAX_D = 768;
constexpr int BOX_BYTES = BN * BK * 2;  // 8 KB: one TMA box = 64 docs x 64 k
constexpr int SLOT_BYTES = 2 * BOX_BYTES;   // a pipeline slot = two boxes (128 k) behind one mbarrier
constexpr int NSLOT = 12;
constexpr int A_COLS = MAX_D / 2;       // 384 TMEM columns hold the query tile
constexpr int TMEM_COLS = 512;
constexpr int NT = 384;                 // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-11 epilogue
constexpr int EPI_THREADS = 256;        // 2 threads per query row (32 accumulator columns each)

constexpr int MODE_FILTER = 0;
constexpr int MODE_STORE = 1;
constexpr int MODE_TILEMAX = 2;   // store only maxima over groups of each thread's 32 columns:
                                  // store[row][(vtile*2 + half) * tm_parts + part]

struct DenseParams {
  const __nv_bfloat16* queries;   // [Q, D]
  int Q, N, D;
  int n_mtiles;                   // ceil(Q / 128)
  int n_vtiles;                   // virtual doc tiles processed by this launch
  int tile_step;                  // real tile = vtile * tile_step (1 = whole corpus, >1 = strided sample)
  int chunk_tiles;                // vtiles per work unit
  int mode;
  int tm_parts;                   // TILEMAX: maxima stored per epilogue thread and tile (1, 2 or 4 groups of 32 / 16 / 8 docs)
  const float* tau;               // [Q]            (FILTER)
  // FILTER output: every epilogue thread owns a private segment per work unit, so survivors are written with
  // plain stores (no atomics, no round trips): segment s = q*n_segs + chunk*2 + half holds seg_cnt[s]
  // (score bits, doc id) pairs, capacity SEG_T; a count > SEG_T flags the query for the exact fallback.
  // Query-major, so the select kernel reads one query's counts and candidates from one contiguous region.
  uint2* seg_pairs;               // [q_pad*n_segs, SEG_T]
  int* seg_cnt;                   // [q_pad*n_segs]
  int q_pad;                      // n_mtiles * 128
  int n_segs;                     // n_chunks * 2
  long long* dbg;                 // optional [gridDim][8] cycle counters of the MMA issuer (profiling builds)
  float* store;                   // [n_mtiles*128, store_ld]  (STORE): column = vtile*64 + j
  long long store_ld;
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u32 bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 addr, u32 parity) {
  u32 done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(u32 smem_dst, const CUtensorMap* tmap, u32 bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<u64>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(u32* smem_dst, u32 ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(u32 taddr, u32 ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem_d] (+)= A[tmem_a] * B[smem desc]   (A from tensor memory, kind::f16 = bf16/f16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_ts(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(u32 bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_st8(u32 taddr, const u32 (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(u32 taddr, u32 (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ u64 make_desc_sw128(u32 saddr) {
  u64 d = 0;
  d |= (u64)((saddr >> 4) & 0x3FFFu);   // start address
  d |= (u64)1 << 16;                    // leading byte offset (ignored for swizzled K-major)
  d |= (u64)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (u64)1 << 46;                    // descriptor version (Blackwell)
  d |= (u64)2 << 61;                    // SWIZZLE_128B
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, N=64, M=128 (one CTA) or 256 (CTA pair)
template <int CG>
struct Idesc {
  static constexpr u32 value = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(BN >> 3) << 17) | ((u32)((BM * CG) >> 4) << 24);
};

constexpr int SEG_T = 96;                 // private candidate segment per (work unit, epilogue thread); round 2: 48 -> 96,
                                          // i.e. twice as long work units on shard-sized corpora (the query tile is
                                          // reloaded into tensor memory once per unit: 31-tile units cost 8 % at 335 k rows)
constexpr int MAX_SLOTS = 24;
constexpr u32 PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-pair peer bit of a shared::cluster address -> CTA rank 0

struct __align__(8) DenseSmemCtl {
  u64 full[MAX_SLOTS];
  u64 empty[MAX_SLOTS];
  u64 tmem_full[2];
  u64 tmem_empty[2];
  u64 a_ready;
  u32 tmem_base;
  u32 pad;
};

__device__ __forceinline__ u32 elect_one() {
  u32 pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ u32 cluster_ctarank() {
  u32 r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(u32 bar, u32 rank) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(u32 smem_dst, const CUtensorMap* tmap, u32 bar, int c0, int c1) {
  // issued by both CTAs of a pair; the transaction bytes are credited to the barrier of CTA rank 0
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<u64>(tmap)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ts_pair(u32 tmem_d, u32 tmem_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(u32 bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(u32* smem_dst, u32 ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(u32 taddr, u32 ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// CG = 1: one CTA per 128-query tile.  CG = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) works on two query
// tiles against the same document tiles: each CTA TMA-loads half of every B tile (32 of the 64 documents), the
// leader CTA issues M=256 MMAs that read both halves, so the L2->SM traffic per FLOP halves.
template <int CG>
__global__ void __launch_bounds__(NT, 1) dense_kernel(const __grid_constant__ CUtensorMap tmap, const DenseParams p) {
  constexpr int BOX_ROWS = BN / CG;                 // documents this CTA loads per tile
  constexpr int BOX_B = BOX_ROWS * BK * 2;          // bytes per TMA box
  constexpr int SLOT_B = 2 * BOX_B;                 // pipeline slot = two boxes (128 k) behind one mbarrier
  constexpr int NSL = 12 * CG;                      // 192 KB ring either way
  constexpr u32 IDESC = Idesc<CG>::value;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  DenseSmemCtl* ctl = reinterpret_cast<DenseSmemCtl*>(smem + NSL * SLOT_B);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const u32 rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int KB2 = (p.D + 2 * BK - 1) / (2 * BK);     // pipeline steps per doc tile (128 k each; last may be 64)
  const int n_chunks = (p.n_vtiles + p.chunk_tiles - 1) / p.chunk_tiles;
  const int n_mgroups = (p.n_mtiles + CG - 1) / CG;
  const int n_units = n_chunks * n_mgroups;
  const int u_first = blockIdx.x / CG, u_step = gridDim.x / CG;
  // shared-window addresses (computed once; barriers are 8 bytes apart)
  const u32 a_slots = smem_u32(smem);
  const u32 a_full = smem_u32(&ctl->full[0]), a_empty = smem_u32(&ctl->empty[0]);
  const u32 a_tfull = smem_u32(&ctl->tmem_full[0]), a_tempty = smem_u32(&ctl->tmem_empty[0]);
  const u32 a_aready = smem_u32(&ctl->a_ready);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < NSL; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->tmem_full[i], 1);
      mbar_init(&ctl->tmem_empty[i], EPI_THREADS * CG);
    }
    mbar_init(&ctl->a_ready, EPI_THREADS * CG);
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) tmem_alloc_pair(&ctl->tmem_base, TMEM_COLS);
    else tmem_alloc(&ctl->tmem_base, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const u32 tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform, one elected lane issues)
    int slot = 0;
    u32 phase = 0;
    for (int u = u_first; u < n_units; u += u_step) {
      const int c = u / n_mgroups;
      const int t0 = c * p.chunk_tiles, t1 = min(p.n_vtiles, t0 + p.chunk_tiles);
      for (int t = t0; t < t1; ++t) {
        const int row0 = t * p.tile_step * BN + (int)rank * BOX_ROWS;
        for (int kb = 0; kb < KB2; ++kb) {
          mbar_wait(a_empty + 8 * slot, phase ^ 1);
          if (elect_one()) {
            const bool two = (kb * 2 + 1) * BK < p.D;
            if (leader) mbar_expect_tx(a_full + 8 * slot, (two ? SLOT_B : BOX_B) * CG);
            const u32 dst = a_slots + slot * SLOT_B;
            if (CG == 2) {
              tma_load_2d_pair(dst, &tmap, a_full + 8 * slot, kb * 2 * BK, row0);
              if (two) tma_load_2d_pair(dst + BOX_B, &tmap, a_full + 8 * slot, (kb * 2 + 1) * BK, row0);
            } else {
              tma_load_2d(dst, &tmap, a_full + 8 * slot, kb * 2 * BK, row0);
              if (two) tma_load_2d(dst + BOX_B, &tmap, a_full + 8 * slot, (kb * 2 + 1) * BK, row0);
            }
          }
          __syncwarp();
          if (++slot == NSL) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if ((warp == 1 || warp == 3) && leader) {
    // ===================================================================== MMA issuers (leader CTA)
    // Issuing one M x 64 x 16 MMA costs the issuing thread ~25 cycles but only 32 tensor-pipe cycles, and every
    // slot adds a barrier wait, so ONE issuer cannot keep the pipe busy (measured: 428 cycles per 256-cycle slot).
    // Two issuer warps alternate document tiles; each owns one accumulator stage (stage == tile parity), its
    // own accumulate chain and its own tcgen05.commit stream, and the tensor pipe interleaves the two chains.
    const int mw = warp == 3 ? 1 : 0;
    int cur_mg = -1;
    u32 a_phase = 0, acc_phase = 0;
    long long n_tile = 0;                        // tiles seen by this CTA so far (all issuers count alike)
    const u64 desc0 = make_desc_sw128(a_slots);
#ifdef QF_DENSE_TIMING
    long long t_full = 0, t_tempty = 0, t_issue = 0, t_aready = 0, n_slots = 0;
    const long long t_begin = clock64();
#define QF_T0 const long long _t0 = clock64();
#define QF_T1(var) var += clock64() - _t0;
#else
#define QF_T0
#define QF_T1(var)
#endif
    for (int u = u_first; u < n_units; u += u_step) {
      const int c = u / n_mgroups, mg = u % n_mgroups;
      const int t0 = c * p.chunk_tiles, t1 = min(p.n_vtiles, t0 + p.chunk_tiles);
      if (mg != cur_mg) {
        { QF_T0 mbar_wait(a_aready, a_phase); QF_T1(t_aready) }
        a_phase ^= 1;
        cur_mg = mg;
        tc_fence_after();
      }
      for (int t = t0; t < t1; ++t, ++n_tile) {
        if ((int)(n_tile & 1) != mw) continue;
        const int acc = mw;
        { QF_T0 mbar_wait(a_tempty + 8 * acc, acc_phase ^ 1); QF_T1(t_tempty) }
        tc_fence_after();
        const u32 tmem_d = tmem_base + A_COLS + acc * BN;
        const long long seq0 = n_tile * KB2;
        for (int kb = 0; kb < KB2; ++kb) {
          const long long seq = seq0 + kb;
          const int slot = (int)(seq % NSL);
          const u32 phase = (u32)((seq / NSL) & 1);
          { QF_T0 mbar_wait(a_full + 8 * slot, phase); QF_T1(t_full) }
          tc_fence_after();
          QF_T0
          if (elect_one()) {
            const u64 desc = desc0 + (u64)(slot * (SLOT_B >> 4));
            const u32 ta = tmem_base + kb * BK;      // 128 k-elements = 64 TMEM columns per step
            const bool two = (kb * 2 + 1) * BK < p.D;
#pragma unroll
            for (int k = 0; k < 4; ++k) {            // +32 B in the B descriptor, +8 columns in the TMEM A operand
              if (CG == 2) umma_ts_pair(tmem_d, ta + k * 8, desc + (u64)(k * 2), IDESC, (kb | k) != 0);
              else umma_ts(tmem_d, ta + k * 8, desc + (u64)(k * 2), IDESC, (kb | k) != 0);
            }
            if (two) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (CG == 2) umma_ts_pair(tmem_d, ta + 32 + k * 8, desc + (u64)((BOX_B >> 4) + k * 2), IDESC, 1u);
                else umma_ts(tmem_d, ta + 32 + k * 8, desc + (u64)((BOX_B >> 4) + k * 2), IDESC, 1u);
              }
            }
            if (CG == 2) {
              umma_commit_pair(a_empty + 8 * slot);
              if (kb == KB2 - 1) umma_commit_pair(a_tfull + 8 * acc);
            } else {
              umma_commit(a_empty + 8 * slot);
              if (kb == KB2 - 1) umma_commit(a_tfull + 8 * acc);
            }
          }
          __syncwarp();
          QF_T1(t_issue)
#ifdef QF_DENSE_TIMING
          ++n_slots;
#endif
        }
        acc_phase ^= 1;
      }
    }
#ifdef QF_DENSE_TIMING
    if (p.dbg && lane == 0 && mw == 0) {
      long long* d = p.dbg + (size_t)blockIdx.x * 8;
      d[0] = clock64() - t_begin; d[1] = t_full; d[2] = t_tempty; d[3] = t_issue; d[4] = t_aready; d[5] = n_slots;
    }
#endif
  } else if (warp >= 4) {
    // ===================================================================== epilogue: 256 threads, 2 per query row
    const int row = ((warp & 3) << 5) | lane;          // TMEM lane == query row of the tile
    const int half = (warp - 4) >> 2;                  // accumulator columns [32*half, 32*half+32)
    const u32 lane_base = (u32)((warp & 3) * 32) << 16;
    int acc = 0, cur_m = -1;
    u32 acc_phase = 0;
    float tau = INFINITY;
    int q = 0;
    for (int u = u_first; u < n_units; u += u_step) {
      const int c = u / n_mgroups, m = (u % n_mgroups) * CG + (int)rank;
      const int t0 = c * p.chunk_tiles, t1 = min(p.n_vtiles, t0 + p.chunk_tiles);
      if (m != cur_m) {
        // every MMA that read the previous query tile has completed (its last tmem_full was consumed)
        q = m * BM + row;
        const bool valid = q < p.Q;
        // this thread loads half of the row's k range: 16-element groups [jb, je)
        const int groups = p.D / 16, jb = half * (groups / 2), je = half ? groups : groups / 2;
        const uint4* src = reinterpret_cast<const uint4*>(p.queries + (size_t)min(q, p.Q - 1) * p.D);
        for (int j = jb; j < je; ++j) {
          uint4 a = __ldg(src + 2 * j), b = __ldg(src + 2 * j + 1);
          u32 r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = 0u;
          }
          tmem_st8(tmem_base + lane_base + j * 8, r);
        }
        tmem_wait_st();
        tc_fence_before();
        if (leader) mbar_arrive(a_aready);
        else mbar_arrive_remote(a_aready, 0);
        cur_m = m;
        tau = (valid && p.mode == MODE_FILTER) ? __ldg(p.tau + q) : INFINITY;
      }
      const size_t seg = (size_t)q * p.n_segs + (size_t)c * 2 + half;
      uint2* seg_out = p.mode == MODE_FILTER ? p.seg_pairs + seg * SEG_T : nullptr;
      int n_mine = 0;
      for (int t = t0; t < t1; ++t) {
        mbar_wait(a_tfull + 8 * acc, acc_phase);
        tc_fence_after();
        u32 v[32];
        tmem_ld32(tmem_base + lane_base + A_COLS + acc * BN + half * 32, v);
        tmem_wait_ld();
        tc_fence_before();
        if (leader) mbar_arrive(a_tempty + 8 * acc);        // accumulator stage is free again
        else mbar_arrive_remote(a_tempty + 8 * acc, 0);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }

        const int doc0 = t * p.tile_step * BN + half * 32;
        if (p.mode == MODE_FILTER) {
          const int nvalid = p.N - doc0;        // >= 32 except in the last tile
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (__uint_as_float(v[j]) >= tau && j < nvalid) {
              if (n_mine < SEG_T) seg_out[n_mine] = make_uint2(v[j], (u32)(doc0 + j));
              ++n_mine;
            }
          }
        } else if (p.mode == MODE_TILEMAX) {
          const int nvalid = p.N - doc0;
          float mq[4];                                    // maxima over the four groups of 8 columns
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 8 * g; j < 8 * g + 8; ++j) mx = (j < nvalid) ? fmaxf(mx, __uint_as_float(v[j])) : mx;
            mq[g] = mx;
          }
          if (m < p.n_mtiles) {
            float* dst = p.store + (size_t)(m * BM + row) * p.store_ld + ((size_t)t * 2 + half) * p.tm_parts;
            if (p.tm_parts == 1) {
              dst[0] = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
            } else if (p.tm_parts == 2) {
              dst[0] = fmaxf(mq[0], mq[1]);
              dst[1] = fmaxf(mq[2], mq[3]);
            } else {
              dst[0] = mq[0]; dst[1] = mq[1]; dst[2] = mq[2]; dst[3] = mq[3];
            }
          }
        } else if (m < p.n_mtiles) {
          float4* dst = reinterpret_cast<float4*>(p.store + (size_t)(m * BM + row) * p.store_ld + (size_t)t * BN + half * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
        }
      }
      if (p.mode == MODE_FILTER && m < p.n_mtiles) p.seg_cnt[seg] = n_mine;      // > SEG_T flags an overflowed segment
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();       // neither CTA may exit (or free TMEM) while its peer still uses it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ threshold
// tau[q] = the j-th largest of the valid sample scores of row q (exact, MSB-first radix select on the
// order-preserving integer image of the floats).  One CTA per query row.
__device__ __forceinline__ u32 ord_f32(float x) {
  const u32 b = __float_as_uint(x);
  return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(u32 o) {
  return __uint_as_float((o >> 31) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(512) tau_kernel(const float* sample, long long ld, int n_vtiles, int tile_step, int N, int Q,
                                                  int target, float* tau, int only_nan) {
  __shared__ u32 hist[256];
  __shared__ u32 s_prefix, s_remaining;
  const int q = blockIdx.x;
  if (q >= Q) return;
  if (only_nan && !isnan(tau[q])) return;
  const float* row = sample + (size_t)q * ld;
  const int n_cols = n_vtiles * BN;
  // number of valid sample docs (real doc index < N)
  long long s_valid = 0;
  for (int t = 0; t < n_vtiles; ++t) {   // cheap: <= 512 iterations, same on every thread
    const long long d0 = (long long)t * tile_step * BN;
    s_valid += max(0ll, min((long long)BN, (long long)N - d0));
  }
  long long jj = ((long long)target * s_valid + N - 1) / N;
  if (jj < 1) jj = 1;
  if (jj > s_valid) jj = s_valid;
  u32 prefix = 0, remaining = (u32)jj;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const u32 mask_hi = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int cidx = threadIdx.x; cidx < n_cols; cidx += blockDim.x) {
      const long long doc = (long long)(cidx / BN) * tile_step * BN + (cidx % BN);
      if (doc < N) {
        const u32 o = ord_f32(row[cidx]);
        if ((o & mask_hi) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 rem = remaining;
      int d = 255;
      for (; d > 0; --d) {
        if (hist[d] >= rem) break;
        rem -= hist[d];
      }
      s_prefix = prefix | ((u32)d << shift);
      s_remaining = rem;
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
  }
  if (threadIdx.x == 0) tau[q] = unord_f32(prefix);
}

// descending bitonic sort of 512 floats in shared memory by 512 threads
__device__ __forceinline__ void bitonic512_desc(float* v, int tid) {
  for (int k = 2; k <= 512; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = tid ^ j;
      if (ixj > tid) {
        const float a = v[tid], b = v[ixj];
        const bool desc = (tid & k) == 0;
        if (desc ? (a < b) : (a > b)) { v[tid] = b; v[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

// Fast path for small order statistics (jj <= 128): the jj-th largest of the per-thread maxima is a lower
// bound L of the answer; the handful of elements >= L are collected and sorted exactly.  Falls back to
// the radix select (tau_kernel) by writing NaN when the collection overflows (host never sees it: the
// caller launches the radix kernel for rows still NaN).
__global__ void __launch_bounds__(512) tau_kernel_small(const float* sample, long long ld, int n_vtiles, int tile_step, int N,
                                                        int Q, int target, float* tau) {
  __shared__ float buf[512];
  __shared__ int s_cnt;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (q >= Q) return;
  const float* row = sample + (size_t)q * ld;
  const int n_cols = n_vtiles * BN;
  long long s_valid = 0;
  for (int t = 0; t < n_vtiles; ++t) {
    const long long d0 = (long long)t * tile_step * BN;
    s_valid += max(0ll, min((long long)BN, (long long)N - d0));
  }
  long long jj = ((long long)target * s_valid + N - 1) / N;
  if (jj < 1) jj = 1;
  if (jj > s_valid) jj = s_valid;
  float mx = -INFINITY;
  for (int cidx = tid; cidx < n_cols; cidx += 512) {
    const long long doc = (long long)(cidx / BN) * tile_step * BN + (cidx % BN);
    if (doc < N) mx = fmaxf(mx, row[cidx]);
  }
  buf[tid] = mx;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  bitonic512_desc(buf, tid);
  const float L = buf[jj - 1];          // at least jj elements are >= L
  __syncthreads();
  for (int cidx = tid; cidx < n_cols; cidx += 512) {
    const long long doc = (long long)(cidx / BN) * tile_step * BN + (cidx % BN);
    if (doc < N) {
      const float v = row[cidx];
      if (v >= L) {
        const int pos = atomicAdd(&s_cnt, 1);
        if (pos < 512) buf[pos] = v;
      }
    }
  }
  __syncthreads();
  const int c = s_cnt;
  if (c > 512) {                        // heavy ties: let the radix kernel handle this row
    if (tid == 0) tau[q] = __int_as_float(0x7fc00000);
    return;
  }
  if (tid >= c) buf[tid] = -INFINITY;
  __syncthreads();
  bitonic512_desc(buf, tid);
  if (tid == 0) tau[q] = buf[jj - 1];
}

// descending bitonic sort of 1024 floats in shared memory by 512 threads
__device__ __forceinline__ void bitonic1024_desc(float* v, int tid) {
  for (int k = 2; k <= 1024; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      // thread t handles the pair (i, i ^ j) with i = the t-th index whose bit j is clear
      const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
      const int ixj = i | j;
      const float a = v[i], b = v[ixj];
      const bool desc = (i & k) == 0;
      if (desc ? (a < b) : (a > b)) { v[i] = b; v[ixj] = a; }
      __syncthreads();
    }
  }
}

// tau[q] = the jj-th largest of the per-(sample tile, column group) maxima written by MODE_TILEMAX: a lower bound
// of the jj-th largest sample score (equal unless two of the top jj fall into the same 32 / 16 / 8-document group), so
// the threshold errs towards admitting a few more candidates.  n_vals = 2 * tm_parts * n_sample_tiles <= 1024.
__global__ void __launch_bounds__(512) tau_kernel_tilemax(const float* vals, long long ld, int n_vals, int jj, int Q, float* tau) {
  __shared__ float buf[1024];
  const int q = blockIdx.x, tid = threadIdx.x;
  if (q >= Q) return;
  const float* row = vals + (size_t)q * ld;
  for (int i = tid; i < 1024; i += 512) buf[i] = i < n_vals ? row[i] : -INFINITY;
  __syncthreads();
  bitonic1024_desc(buf, tid);
  if (tid == 0) tau[q] = buf[jj - 1];
}

__global__ void fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------ select
// Rank one query's candidates by (score desc, id asc) and emit the top k.  Candidates come either from
// the FILTER pass (n = cand_cnt[q], arbitrary order) or from `nseg` already-ranked segments
// (multi-GPU merge).  Stable LSD radix on the 32-bit score image, exact fix-up of equal-score runs by id.
constexpr int SEL_CAP = 8192;
constexpr int SEL_NT = 512;
constexpr int SEL_ROUNDS = (SEL_CAP / 32 + SEL_NT / 32 - 1) / (SEL_NT / 32);

struct SelectParams {
  // source A (FILTER pass): per-(unit, thread) private segments, see DenseParams
  const uint2* seg_pairs;
  const int* seg_cnt;
  int n_segs;
  // source B (multi-GPU merge): nseg already-ranked segments of seg_len entries
  const float* cand_score;
  const int* cand_id;
  long long q_stride;      // element stride between queries
  long long seg_stride;    // element stride between segments
  int nseg, seg_len;
  int Q, k, N_total;
  long long id_offset;
  float* out_scores;       // [Q, k]
  int* out_ids;            // [Q, k]  (-1 padded)
  int* out_status;         // [Q] nullable: 1 = candidate underflow, 2 = overflow
  int fixup;               // select_kernel only: process just the queries select2_kernel deferred
  // output addressing: row(q) = (q / chunk_rows) * chunk_stride + (q % chunk_rows) * out_stride  (elements); the
  // plain layout is out_stride = k, chunk_rows = 0.  The packed layout of the multi-GPU exchange interleaves score
  // and id rows (out_stride = 2k, out_ids = out_scores + k) and leaves one spare row per destination chunk.
  long long out_stride;
  int chunk_rows;
  long long chunk_stride;
};
__device__ __forceinline__ size_t out_row(const SelectParams& p, int q) {
  if (p.chunk_rows > 0) return (size_t)(q / p.chunk_rows) * (size_t)p.chunk_stride + (size_t)(q % p.chunk_rows) * (size_t)p.out_stride;
  return (size_t)q * (size_t)p.out_stride;
}
constexpr int SEL_DEFERRED = -2;   // out_ids[q][0] marker: query left to select_kernel by select2_kernel

struct SelSmem {
  float* score;
  int* id;
  u32 *keyA, *keyB;
  u16 *idxA, *idxB, *whist;
};

__device__ __forceinline__ SelSmem carve_sel(unsigned char* sm) {
  SelSmem s;
  s.score = reinterpret_cast<float*>(sm);
  s.id = reinterpret_cast<int*>(sm + 4 * SEL_CAP);
  s.keyA = reinterpret_cast<u32*>(sm + 8 * SEL_CAP);
  s.keyB = reinterpret_cast<u32*>(sm + 12 * SEL_CAP);
  s.idxA = reinterpret_cast<u16*>(sm + 16 * SEL_CAP);
  s.idxB = reinterpret_cast<u16*>(sm + 18 * SEL_CAP);
  s.whist = reinterpret_cast<u16*>(sm + 20 * SEL_CAP);
  return s;
}

// Keep only candidates that can be among the `keep` best (score desc): MSB-first radix select of the keep-th
// largest score on the order-preserving integer image, then an in-smem compaction of everything >= it
// (all ties of the boundary score are kept, so exactness is untouched).  Returns the new count.
__device__ int prefilter_topk(const SelSmem& S, int n, int keep, int* s_scratch /* >= 260 ints */) {
  const int tid = threadIdx.x;
  if (n <= keep + keep / 2) return n;
  int* hist = s_scratch;            // [256]
  int* ctl = s_scratch + 256;       // [0] prefix, [1] remaining, [2] out count
  u32 prefix = 0;
  int remaining = keep;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const u32 mask_hi = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < 256; i += SEL_NT) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_NT) {
      const u32 key = ~ord_f32(S.score[i] + 0.0f);            // ascending key = descending score
      if ((key & mask_hi) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < 32) {
      // warp 0: find the digit whose cumulative count reaches `remaining`
      int c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[tid * 8 + j]; tot += c[j]; }
      int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t_ = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += t_;
      }
      int run = incl - tot;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (run < remaining && run + c[j] >= remaining) { ctl[0] = (int)(prefix | ((u32)(tid * 8 + j) << shift)); ctl[1] = remaining - run; }
        run += c[j];
      }
    }
    __syncthreads();
    prefix = (u32)ctl[0];
    remaining = ctl[1];
    __syncthreads();
  }
  // compaction into (keyA: score bits, keyB: id), then back
  if (tid == 0) ctl[2] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += SEL_NT) {
    const float sc = S.score[i];
    if (~ord_f32(sc + 0.0f) <= prefix) {
      const int pos = atomicAdd(&ctl[2], 1);
      S.keyA[pos] = __float_as_uint(sc);
      S.keyB[pos] = (u32)S.id[i];
    }
  }
  __syncthreads();
  const int m = ctl[2];
  const int m_pad = (m + 31) & ~31;
  for (int i = tid; i < m_pad; i += SEL_NT) {
    S.score[i] = i < m ? __uint_as_float(S.keyA[i]) : -INFINITY;
    S.id[i] = i < m ? (int)S.keyB[i] : 0x7fffffff;
  }
  __syncthreads();
  return m;
}

// Rank the n candidates held in (S.score, S.id) by (score desc, id asc).  Block-wide; returns the index
// array holding the order.  Entries [n, n_pad) must have been filled with score -inf / id INT_MAX.
__device__ const u16* rank_candidates(const SelSmem& S, int n, int* warp_tot, int* flag_single, int* flag_long) {
  const int tid = threadIdx.x;
  const int n_pad = (n + 31) & ~31;
  if (tid == 0) *flag_long = 0;
  for (int i = tid; i < n_pad; i += SEL_NT) {
    S.keyA[i] = i < n ? ~ord_f32(S.score[i] + 0.0f) : 0xFFFFFFFFu;     // ascending key = descending score
    S.idxA[i] = (u16)i;
  }
  __syncthreads();
  u32* kin = S.keyA; u32* kout = S.keyB;
  u16* iin = S.idxA; u16* iout = S.idxB;
#pragma unroll 1
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    const u32* kk = kin;
    auto dig = [&](int e) -> u32 { return (kk[e] >> shift) & 255u; };
    if (qf::radix_pass<SEL_NT, SEL_ROUNDS, true>(dig, n_pad, kin, kout, iin, iout, S.whist, warp_tot, flag_single)) {
      u32* t = kin; kin = kout; kout = t;
      u16* ti = iin; iin = iout; iout = ti;
    }
  }
  // equal-score runs: order by id ascending (exact); long runs take the two-key radix path below
  for (int j = tid; j < n; j += SEL_NT) {
    const u32 kj = kin[j];
    const bool head = (j == 0 || kin[j - 1] != kj) && (j + 1 < n && kin[j + 1] == kj);
    if (head) {
      int e = j + 2;
      while (e < n && kin[e] == kj) ++e;
      if (e - j > 48) {
        *flag_long = 1;
      } else {
        for (int a = j + 1; a < e; ++a) {
          const u16 ia = iin[a];
          const int ida = S.id[ia];
          int b = a - 1;
          while (b >= j && S.id[iin[b]] > ida) { iin[b + 1] = iin[b]; --b; }
          iin[b + 1] = ia;
        }
      }
    }
  }
  __syncthreads();
  if (*flag_long) {
    // many equal scores: LSD radix over (id, then score) = full (score desc, id asc) order, no fix-up needed
    for (int i = tid; i < n_pad; i += SEL_NT) S.idxA[i] = (u16)i;
    __syncthreads();
    iin = S.idxA; iout = S.idxB;
#pragma unroll 1
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 8 * (pass & 3);
      const u16* ii = iin;
      const bool by_id = pass < 4;
      auto dig = [&](int e) -> u32 {
        const int c = ii[e];
        const u32 key = by_id ? (u32)S.id[c] : (c < n ? ~ord_f32(S.score[c] + 0.0f) : 0xFFFFFFFFu);
        return (key >> shift) & 255u;
      };
      if (qf::radix_pass<SEL_NT, SEL_ROUNDS, false>(dig, n_pad, nullptr, nullptr, iin, iout, S.whist, warp_tot, flag_single)) {
        u16* ti = iin; iin = iout; iout = ti;
      }
    }
  }
  return iin;
}

__global__ void __launch_bounds__(SEL_NT, 1) select_kernel(const SelectParams p) {
  extern __shared__ __align__(16) unsigned char sm[];
  const SelSmem S = carve_sel(sm);
  __shared__ int warp_tot[32];
  __shared__ int flag_single, flag_long;
  const int tid = threadIdx.x;

  __shared__ int s_n, s_flag;
  __shared__ int s_pref[260];
  for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
    if (p.fixup && p.out_ids[out_row(p, q)] != SEL_DEFERRED) continue;      // uniform: same word for every thread
    int status = 0;
    int n = 0;
    const int want = min(p.k, p.N_total);
    if (p.seg_pairs) {
      // gather the query's private segments: counts first (contiguous), block scan -> offsets, then all
      // segments are copied with independent loads
      if (tid == 0) { s_n = 0; s_flag = 0; }
      int* s_off = reinterpret_cast<int*>(S.keyA);                 // [n_segs + 1] scratch (sort buffers are idle)
      const size_t seg0 = (size_t)q * p.n_segs;
      __syncthreads();
      {
        // exclusive scan of min(cnt, SEG_T) over the n_segs segments, SEL_NT at a time
        int carry = 0;
        for (int base = 0; base < p.n_segs; base += SEL_NT) {
          const int sg = base + tid;
          int c = sg < p.n_segs ? p.seg_cnt[seg0 + sg] : 0;
          if (c > SEG_T) { c = SEG_T; s_flag = 2; }                 // a unit overflowed its private segment
          int incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t_ = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += t_;
          }
          if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
          __syncthreads();
          int wb = 0;
          for (int w = 0; w < (tid >> 5); ++w) wb += warp_tot[w];
          if (sg < p.n_segs) s_off[sg] = carry + wb + incl - c;
          int tot = 0;
          for (int w = 0; w < SEL_NT / 32; ++w) tot += warp_tot[w];
          carry += tot;
          __syncthreads();
        }
        if (tid == 0) { s_off[p.n_segs] = carry; s_n = carry; }
      }
      __syncthreads();
      {
        // flat gather: entry e -> its segment by binary search over the offsets, all loads independent
        const int total = min(s_off[p.n_segs], SEL_CAP);
        for (int e = tid; e < total; e += SEL_NT) {
          int lo = 0, hi = p.n_segs;                    // largest sg with s_off[sg] <= e
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= e) lo = mid; else hi = mid;
          }
          const uint2 pr = p.seg_pairs[(seg0 + lo) * SEG_T + (e - s_off[lo])];
          S.score[e] = __uint_as_float(pr.x);
          S.id[e] = (int)pr.y;
        }
      }
      __syncthreads();
      n = s_n;
      status |= s_flag;
      if (n > SEL_CAP) { status |= 2; n = SEL_CAP; }
      if (n < want) status |= 1;
      const int n_pad = (n + 31) & ~31;
      for (int i = n + tid; i < n_pad; i += SEL_NT) { S.score[i] = -INFINITY; S.id[i] = 0x7fffffff; }
    } else {
      n = p.nseg * p.seg_len;
      if (n > SEL_CAP) { status |= 2; n = SEL_CAP; }
      const int n_pad = (n + 31) & ~31;
      for (int i = tid; i < n_pad; i += SEL_NT) {
        float s = -INFINITY;
        int id = 0x7fffffff;
        if (i < n) {
          const size_t a = (size_t)q * p.q_stride + (size_t)(i / p.seg_len) * p.seg_stride + (i % p.seg_len);
          s = p.cand_score[a];
          id = p.cand_id[a];
          if (id < 0) { s = -INFINITY; id = 0x7fffffff; }         // padding of a short segment
        }
        S.score[i] = s;
        S.id[i] = id;
      }
    }
    __syncthreads();
    n = prefilter_topk(S, n, p.k, s_pref);
    const u16* order = rank_candidates(S, n, warp_tot, &flag_single, &flag_long);
    for (int j = tid; j < p.k; j += SEL_NT) {
      float s = -INFINITY;
      int id = -1;
      if (j < n) {
        const int e = order[j];
        if (S.id[e] != 0x7fffffff && S.score[e] != -INFINITY) {
          s = S.score[e];
          id = (int)(S.id[e] + p.id_offset);
        }
      }
      p.out_scores[out_row(p, q) + j] = s;
      p.out_ids[out_row(p, q) + j] = id;
    }
    if (p.out_status && tid == 0) p.out_status[q] = status;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// select2_kernel: the common-case candidate ranking after the FILTER pass -- several queries in flight per SM and ONE
// bucket pass instead of a 4-pass radix select followed by a 4-pass radix sort (same idea as the fusion kernel's
// ranking, csrc/qf_fuse_stream.cu).  Per query (one CTA of 256 threads, ~75 KB of shared memory, 3 CTAs per SM):
//   1. scan the per-(unit, thread) segment counts, copy the ~4k (score, id) pairs into shared memory
//      (thread per segment; neighbouring threads read neighbouring segments, L1 absorbs the sector re-use);
//   2. a 256-candidate sample gives min / max and a 64-bin histogram -> monotone piecewise-linear map
//      score -> one of 4096 buckets (histogram equalisation), higher score = lower bucket;
//   3. one pass counts the buckets (packed 16-bit shared-memory atomics); a block scan finds the first bucket b*
//      whose cumulative count reaches k: only candidates in buckets <= b* can be in the top k;
//   4. those (k .. k + one bucket) candidates are scattered into bucket order and each one's exact rank is its
//      bucket's start plus the number of bucket mates that precede it under (score desc, id asc).
// Anything unusual -- more than S2_CAP candidates, a segment overflow, > 1024 segments, a bucket longer than
// S2_LONGB inside the cut (long runs of equal scores) -- marks the query SEL_DEFERRED for select_kernel (fix-up mode).
constexpr int S2_NT = 256;
constexpr int S2_CAP = 6144;        // candidates held per query after the FILTER pass (3 CTAs per SM)
constexpr int S2_CAP_MERGE = 8192;  // candidates held per query in the k-way merge of qf_topk_merge (2 CTAs per SM)
constexpr int S2_NB = 4096;         // buckets
constexpr int S2_KEEP = 2048;       // candidates that survive the bucket cut (k + one bucket)
constexpr int S2_MAXSEG = 1024;
constexpr int S2_LONGB = 128;
constexpr int S2_CB = 64;
template <int CAP>
struct S2Layout {
  static constexpr size_t off_id = 4 * (size_t)CAP;
  static constexpr size_t off_bc = 8 * (size_t)CAP;
  static constexpr size_t off_cnt = 10 * (size_t)CAP;
  static constexpr size_t off_order = off_cnt + 2 * S2_NB;        // u16 order[S2_KEEP]; aliases u16 seg_off[S2_MAXSEG + 1]
  static constexpr size_t off_coarse = off_order + 2 * S2_KEEP;
  static constexpr size_t total = off_coarse + 4 * S2_CB * 3;
};
static_assert(2 * (S2_MAXSEG + 1) <= 2 * S2_KEEP && S2_NB % (S2_NT * 8) == 0 && S2_CAP_MERGE < 65536, "select2 layout");

template <int CAP, int MINB>
__global__ void __launch_bounds__(S2_NT, MINB) select2_kernel(const SelectParams p) {
  using L2 = S2Layout<CAP>;
  extern __shared__ __align__(16) unsigned char sm[];
  float* s_score = reinterpret_cast<float*>(sm);
  int* s_id = reinterpret_cast<int*>(sm + L2::off_id);
  u16* s_bc = reinterpret_cast<u16*>(sm + L2::off_bc);
  u32* s_cnt32 = reinterpret_cast<u32*>(sm + L2::off_cnt);
  u16* s_cur16 = reinterpret_cast<u16*>(sm + L2::off_cnt);
  u16* s_order = reinterpret_cast<u16*>(sm + L2::off_order);
  u16* s_segoff = reinterpret_cast<u16*>(sm + L2::off_order);
  int* s_chist = reinterpret_cast<int*>(sm + L2::off_coarse);
  float* s_ccum = reinterpret_cast<float*>(sm + L2::off_coarse + 4 * S2_CB);
  float* s_cw = reinterpret_cast<float*>(sm + L2::off_coarse + 8 * S2_CB);
  __shared__ int warp_tot[S2_NT / 32];
  __shared__ float red_mn[S2_NT / 32], red_mx[S2_NT / 32];
  __shared__ int s_bstar, s_kept;
  constexpr int NW = S2_NT / 32, CPT = S2_NB / S2_NT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int q = blockIdx.x; q < p.Q; q += gridDim.x) {
    __syncthreads();                               // previous query's shared state is dead
    const int want_k = min(p.k, p.N_total);
    const size_t seg0 = (size_t)q * p.n_segs;
    if (tid == 0) { s_bstar = S2_NB - 1; s_kept = 0; }
    for (int i = tid; i < S2_NB / 2; i += S2_NT) s_cnt32[i] = 0u;
    for (int i = tid; i < S2_CB; i += S2_NT) s_chist[i] = 0;
    int n = 0, n_valid = 0;
    if (p.seg_pairs) {
      // ---- 1. segment counts -> offsets
      int my_c[S2_MAXSEG / S2_NT];
      int carry = 0;
      bool bad = p.n_segs > S2_MAXSEG;
#pragma unroll
      for (int r = 0; r < S2_MAXSEG / S2_NT; ++r) {
        const int sg = r * S2_NT + tid;
        int c = (!bad && sg < p.n_segs) ? __ldg(p.seg_cnt + seg0 + sg) : 0;
        if (c > SEG_T) { bad = true; c = SEG_T; }    // a unit overflowed its private segment: select_kernel reports it
        my_c[r] = c;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t_ = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t_;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int wb = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          if (w < warp) wb += warp_tot[w];
          tot += warp_tot[w];
        }
        s_segoff[sg] = (u16)min(carry + wb + incl - c, 65535);
        carry += tot;
        __syncthreads();
      }
      n = n_valid = carry;
      if (__syncthreads_or((bad || n > CAP) ? 1 : 0)) {
        if (tid == 0) p.out_ids[out_row(p, q)] = SEL_DEFERRED;
        continue;
      }
      // ---- copy the candidates (thread per segment)
#pragma unroll
      for (int r = 0; r < S2_MAXSEG / S2_NT; ++r) {
        const int sg = r * S2_NT + tid;
        const int c = my_c[r];
        if (c > 0) {
          const uint2* src = p.seg_pairs + (seg0 + sg) * SEG_T;
          const int off = s_segoff[sg];
          for (int e = 0; e < c; ++e) {
            const uint2 pr = __ldg(src + e);
            s_score[off + e] = __uint_as_float(pr.x);
            s_id[off + e] = (int)pr.y;
          }
        }
      }
    } else {
      // ---- k-way merge input: nseg ranked lists of seg_len entries (id < 0 = padding of a short list)
      n = p.nseg * p.seg_len;
      if (n > CAP) {                                   // uniform
        if (tid == 0) p.out_ids[out_row(p, q)] = SEL_DEFERRED;
        continue;
      }
      int valid = 0;
      for (int i = tid; i < n; i += S2_NT) {
        const size_t a = (size_t)q * p.q_stride + (size_t)(i / p.seg_len) * p.seg_stride + (i % p.seg_len);
        float v = __ldg(p.cand_score + a);
        int id = __ldg(p.cand_id + a);
        if (id < 0 || v == -INFINITY) { v = -INFINITY; id = 0x7fffffff; } else ++valid;
        s_score[i] = v;
        s_id[i] = id;
      }
      valid = __reduce_add_sync(0xffffffffu, valid);
      if (lane == 0) warp_tot[warp] = valid;
      __syncthreads();
      n_valid = 0;
#pragma unroll
      for (int w = 0; w < NW; ++w) n_valid += warp_tot[w];
    }
    __syncthreads();
    const int want = min(want_k, n_valid);
    // ---- 2. sample -> range, equalisation map (padding entries, score -inf, always go to the last bucket)
    float mn = INFINITY, mx = -INFINITY, samp = 0.f;
    bool has = tid < n;
    if (has) {
      samp = s_score[(int)(((long long)tid * n) / S2_NT)];     // spread over the whole candidate array
      has = samp != -INFINITY;
      if (has) mn = mx = samp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red_mn[warp] = mn; red_mx[warp] = mx; }
    __syncthreads();
    mn = INFINITY; mx = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) { mn = fminf(mn, red_mn[w]); mx = fmaxf(mx, red_mx[w]); }
    float cscale = (mx > mn) ? (float)S2_CB / (mx - mn) : 0.f;
    if (!(cscale > 0.f) || isinf(cscale)) cscale = 0.f;
    auto coarse = [&](float v, float& frac) -> int {
      float tf = (mx - v) * cscale;                              // monotone: higher score -> smaller tf
      tf = fminf(fmaxf(tf, 0.f), (float)S2_CB);
      int c = (int)tf;
      c = c > S2_CB - 1 ? S2_CB - 1 : c;
      frac = tf - (float)c;
      return c;
    };
    if (has) { float fr; atomicAdd(&s_chist[coarse(samp, fr)], 1); }
    __syncthreads();
    if (warp == 0) {
      const int h0 = s_chist[2 * lane] + 1, h1 = s_chist[2 * lane + 1] + 1;
      const int tot = __reduce_add_sync(0xffffffffu, h0 + h1);
      const int w0 = (h0 * S2_NB) / tot, w1 = (h1 * S2_NB) / tot;
      int incl = w0 + w1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t_ = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t_;
      }
      const int excl = incl - (w0 + w1);
      s_ccum[2 * lane] = (float)excl;
      s_cw[2 * lane] = (float)w0;
      s_ccum[2 * lane + 1] = (float)(excl + w0);
      s_cw[2 * lane + 1] = (float)w1;
    }
    __syncthreads();
    // ---- 3. bucket counts
    for (int i = tid; i < n; i += S2_NT) {
      float frac;
      const float v = s_score[i];
      const int c = coarse(v, frac);
      int b = (int)fmaf(frac, s_cw[c], s_ccum[c]);
      b = (b < 0 ? 0 : (b > S2_NB - 1 ? S2_NB - 1 : b));
      if (v == -INFINITY) b = S2_NB - 1;
      s_bc[i] = (u16)b;
      atomicAdd(&s_cnt32[b >> 1], 1u << ((b & 1) * 16));
    }
    __syncthreads();
    {   // exclusive scan of the 16-bit counters (CPT consecutive per thread), in place; cut bucket b*
      int tsum = 0;
      uint4 cv[CPT / 8];
      int cnts[CPT];
#pragma unroll
      for (int v = 0; v < CPT / 8; ++v) {
        cv[v] = reinterpret_cast<const uint4*>(s_cnt32)[tid * (CPT / 8) + v];
        u32 wds[4] = {cv[v].x, cv[v].y, cv[v].z, cv[v].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c0 = (int)(wds[i] & 0xFFFFu), c1 = (int)(wds[i] >> 16);
          cnts[v * 8 + 2 * i] = c0;
          cnts[v * 8 + 2 * i + 1] = c1;
          wds[i] = (u32)tsum | ((u32)(tsum + c0) << 16);
          tsum += c0 + c1;
        }
        cv[v] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
      }
      int incl = tsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t_ = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t_;
      }
      if (lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      int wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
      const int base = wbase + incl - tsum;
      // buckets of this thread: the cut, and long buckets inside the cut
      int run = base, tmax = 0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        if (run < want) {                              // bucket starts inside the top `want`
          tmax = cnts[i] > tmax ? cnts[i] : tmax;
          if (run + cnts[i] >= want) { s_bstar = tid * CPT + i; s_kept = run + cnts[i]; }
        }
        run += cnts[i];
      }
      const u32 base2 = (u32)base * 0x00010001u;
#pragma unroll
      for (int v = 0; v < CPT / 8; ++v) {
        cv[v].x += base2; cv[v].y += base2; cv[v].z += base2; cv[v].w += base2;
        reinterpret_cast<uint4*>(s_cnt32)[tid * (CPT / 8) + v] = cv[v];
      }
      if (__syncthreads_or(tmax > S2_LONGB ? 1 : 0)) {
        if (tid == 0) p.out_ids[out_row(p, q)] = SEL_DEFERRED;
        continue;
      }
    }
    const int bstar = s_bstar, kept = want > 0 ? s_kept : 0;
    if (kept > S2_KEEP) {                               // uniform; cannot happen with S2_LONGB < S2_KEEP - k_max... guard
      if (tid == 0) p.out_ids[out_row(p, q)] = SEL_DEFERRED;
      continue;
    }
    // ---- 4. scatter the survivors into bucket order (cursor = bucket base, advanced atomically)
    for (int i = tid; i < n; i += S2_NT) {
      const int b = s_bc[i];
      if (b <= bstar && want > 0) {
        const int sh = (b & 1) * 16;
        const u32 old = atomicAdd(&s_cnt32[b >> 1], 1u << sh);
        s_order[(old >> sh) & 0xFFFFu] = (u16)i;
      }
    }
    __syncthreads();
    // ---- exact rank inside the bucket, output
    float* osc = p.out_scores + out_row(p, q);
    int* oid = p.out_ids + out_row(p, q);
    for (int j = tid; j < kept; j += S2_NT) {
      const int i = s_order[j];
      const float v = s_score[i];
      const int id = s_id[i];
      const int b = s_bc[i];
      const int hi = s_cur16[b];                        // after the scatter: end of bucket b = start of bucket b + 1
      const int lo = b > 0 ? (int)s_cur16[b - 1] : 0;
      int rank = lo;
      for (int m = lo; m < hi; ++m) {
        const int i2 = s_order[m];
        const float v2 = s_score[i2];
        rank += (v2 > v || (v2 == v && s_id[i2] < id)) ? 1 : 0;
      }
      if (rank < want) {
        osc[rank] = v;
        oid[rank] = (int)(id + p.id_offset);           // rank < want <= n_valid: never a padding entry
      }
    }
    for (int j = want + tid; j < p.k; j += S2_NT) {
      osc[j] = -INFINITY;
      oid[j] = -1;
    }
    if (p.out_status && tid == 0) p.out_status[q] = n_valid < want_k ? 1 : 0;
  }
}

// Exact streaming top-k of stored score rows (the fallback when the sampled threshold admitted too few or too
// many candidates for a query).  One CTA per row: candidates >= the running k-th score are appended to the
// shared-memory buffer; when it fills up it is ranked and cut back to k.  Documents are scanned in increasing
// id order, so once k candidates are held a later document with score == k-th can never enter (its id is
// larger): strict '>' keeps the result exact under arbitrary ties.
struct RowTopkParams {
  const float* scores;     // [n_rows, ld]
  long long ld;
  const int* rows;         // [n_rows] output row (query index) of each stored row
  int n_rows, N, k;
  long long id_offset;
  float* out_scores;       // [Q, k]
  int* out_ids;
  int* out_status;         // nullable; cleared to 0 for repaired rows
};

__global__ void __launch_bounds__(SEL_NT, 1) row_topk_kernel(const RowTopkParams p) {
  extern __shared__ __align__(16) unsigned char sm[];
  const SelSmem S = carve_sel(sm);
  __shared__ int warp_tot[32];
  __shared__ int flag_single, flag_long, s_cnt;
  const int tid = threadIdx.x;
  constexpr int BATCH = SEL_NT * 4;
  for (int r = blockIdx.x; r < p.n_rows; r += gridDim.x) {
    const float* row = p.scores + (size_t)r * p.ld;
    const int q = p.rows[r];
    float thr = -INFINITY;
    bool have_k = false;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int base = 0; base < p.N; base += BATCH) {
      if (s_cnt + BATCH > SEL_CAP) {           // make room: rank, keep the best k
        const int n = s_cnt;
        const int n_pad = (n + 31) & ~31;
        for (int i = n + tid; i < n_pad; i += SEL_NT) { S.score[i] = -INFINITY; S.id[i] = 0x7fffffff; }
        __syncthreads();
        const u16* order = rank_candidates(S, n, warp_tot, &flag_single, &flag_long);
        const int keep = min(n, p.k);
        float ks[(SEL_CAP / SEL_NT) > 2 ? 2 : 2];
        int ki[2];
        // k <= 1024 = 2 * SEL_NT: each thread carries at most two kept entries through the compaction
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int j = tid + t * SEL_NT;
          if (j < keep) { ks[t] = S.score[order[j]]; ki[t] = S.id[order[j]]; }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int j = tid + t * SEL_NT;
          if (j < keep) { S.score[j] = ks[t]; S.id[j] = ki[t]; }
        }
        if (tid == 0) s_cnt = keep;
        __syncthreads();
        have_k = keep >= p.k;
        thr = have_k ? S.score[keep - 1] : -INFINITY;
        __syncthreads();
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int d = base + t * SEL_NT + tid;
        if (d < p.N) {
          const float s = row[d];
          if (have_k ? (s > thr) : (s >= thr)) {
            const int pos = atomicAdd(&s_cnt, 1);
            S.score[pos] = s;
            S.id[pos] = d;
          }
        }
      }
      __syncthreads();
    }
    const int n = s_cnt;
    const int n_pad = (n + 31) & ~31;
    for (int i = n + tid; i < n_pad; i += SEL_NT) { S.score[i] = -INFINITY; S.id[i] = 0x7fffffff; }
    __syncthreads();
    const u16* order = rank_candidates(S, n, warp_tot, &flag_single, &flag_long);
    for (int j = tid; j < p.k; j += SEL_NT) {
      float s = -INFINITY;
      int id = -1;
      if (j < n) { s = S.score[order[j]]; id = (int)(S.id[order[j]] + p.id_offset); }
      p.out_scores[(size_t)q * p.k + j] = s;
      p.out_ids[(size_t)q * p.k + j] = id;
    }
    if (p.out_status && tid == 0) p.out_status[q] = 0;
    __syncthreads();
  }
}

constexpr size_t SEL_SMEM = 20 * SEL_CAP + 2 * (SEL_NT / 32) * 256;

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    QF_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    QF_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, QF_E_CUDA, "cuTensorMapEncodeTiled not available");
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return QF_OK;
}

int make_corpus_tmap(const void* corpus, int N, int D, int box_rows, CUtensorMap* tmap) {
  EncodeTiledFn encode;
  int rc = get_encode_fn(&encode);
  if (rc != QF_OK) return rc;
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstride[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(corpus), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QF_REQUIRE(r == CUDA_SUCCESS, QF_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return QF_OK;
}

struct DenseMaps {
  CUtensorMap one;    // 64-document boxes (single CTA)
  CUtensorMap pair;   // 32-document boxes (each CTA of a pair loads half a tile)
};

constexpr size_t DENSE_SMEM_BYTES = (size_t)12 * 16384 + 512 + 1024;

// Launch the scoring kernel: CTA pairs (cta_group::2) when there are at least two query tiles, else single CTAs.
int launch_dense(const DenseMaps& maps, const DenseParams& p_in, cudaStream_t st) {
  DenseParams p = p_in;
#ifdef QF_DENSE_TIMING
  static long long* dbg = nullptr;
  if (!dbg) cudaMalloc(&dbg, sizeof(long long) * 8 * 256);
  p.dbg = dbg;
#endif
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[qf::device_index()];   // kernel attributes are per device
  if (!configured) {
    QF_CUDA_TRY(cudaFuncSetAttribute(dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DENSE_SMEM_BYTES));
    QF_CUDA_TRY(cudaFuncSetAttribute(dense_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DENSE_SMEM_BYTES));
    configured = true;
  }
  const int sms = qf::sm_count();
  const int n_chunks = (p.n_vtiles + p.chunk_tiles - 1) / p.chunk_tiles;
  const bool force1 = getenv("QF_DENSE_SINGLE_CTA") != nullptr;
  if (p.n_mtiles >= 2 && !force1) {
    const int units = n_chunks * ((p.n_mtiles + 1) / 2);
    int clusters = sms / 2;
    if (clusters > units) clusters = units;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = DENSE_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QF_CUDA_TRY(cudaLaunchKernelEx(&cfg, dense_kernel<2>, maps.pair, p));
  } else {
    const int units = n_chunks * p.n_mtiles;
    dense_kernel<1><<<units < sms ? units : sms, NT, DENSE_SMEM_BYTES, st>>>(maps.one, p);
  }
  QF_CUDA_TRY(cudaGetLastError());
#ifdef QF_DENSE_TIMING
  if (p.mode == MODE_FILTER && p.n_vtiles > 10000) {
    static long long h[8 * 256];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
    for (int b = 0; b < 4; ++b)
      fprintf(stderr, "[dense timing] cta %d: total %lld  wait_full %lld  wait_tmem_empty %lld  issue %lld  wait_a %lld  slots %lld\n",
              b * 2, h[b * 16], h[b * 16 + 1], h[b * 16 + 2], h[b * 16 + 3], h[b * 16 + 4], h[b * 16 + 5]);
  }
#endif
  return QF_OK;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct DenseWs {
  size_t off_tau, off_segcnt, off_segpairs, off_sample, total;
  int n_sample_tiles, chunk_tiles, n_chunks;
  long long sample_ld;
};

// main-pass chunking: with a sampled threshold a thread sees ~12 survivors per 256-tile unit (units are sized for
// SEG_T / 4 expected survivors, i.e. 4x head-room; overflow is flagged and repaired exactly); without a threshold (tiny corpora, tau = -inf) every
// document survives, so units are single tiles (32 documents per epilogue thread <= SEG_T).
int cand_target(int k) { return 3 * k + 1000; }   // candidates the sampled threshold aims to admit per query

int main_chunk_tiles(int N, int k) {
  if (N <= SEL_CAP) return 1;
  // expected survivors per epilogue thread and unit = chunk_tiles * 32 docs * target / N  <=  SEG_T / 4
  long long c = (long long)(SEG_T / 4) * N / (32LL * cand_target(k));
  if (c < 1) c = 1;
  if (c > 256) c = 256;
  return (int)c;
}

DenseWs plan_ws(int Q, int N, int k) {
  DenseWs w{};
  const int n_mtiles = (Q + BM - 1) / BM;
  const long long T = ((long long)N + BN - 1) / BN;
  // sample size: at most 512 tiles, and small enough that the order statistic jj = target * S / N stays <= 64 so
  // the cheap tile-max estimator applies (S <= 64 N / target documents, i.e. N / target tiles)
  long long ns = (long long)N / cand_target(k);
  if (ns < 8) ns = 8;
  if (ns > 512) ns = 512;
  if (ns > T) ns = T;
  w.n_sample_tiles = N > SEL_CAP ? (int)ns : 0;
  w.sample_ld = (long long)w.n_sample_tiles * BN;
  w.chunk_tiles = main_chunk_tiles(N, k);
  w.n_chunks = (int)((T + w.chunk_tiles - 1) / w.chunk_tiles);
  const size_t n_seg = (size_t)w.n_chunks * 2 * n_mtiles * BM;
  size_t o = 0;
  w.off_tau = o;      o = align_up(o + sizeof(float) * (size_t)Q, 256);
  w.off_segcnt = o;   o = align_up(o + sizeof(int) * n_seg, 256);
  w.off_segpairs = o; o = align_up(o + sizeof(uint2) * n_seg * SEG_T, 256);
  w.off_sample = o;   o = align_up(o + sizeof(float) * (size_t)n_mtiles * BM * (size_t)w.sample_ld, 256);
  w.total = o;
  return w;
}

constexpr size_t DENSE_SMEM = (size_t)12 * 16384 + sizeof(DenseSmemCtl) + 1024;

}  // namespace

extern "C" size_t qf_dense_workspace_bytes(int Q, int N, int k) {
  if (Q < 1 || N < 1 || k < 1) return 0;
  return plan_ws(Q, N, k).total;
}

extern "C" int qf_dense_max_k(void) { return 1024; }

namespace {
struct OutLayout {           // see SelectParams: row(q) = (q / chunk_rows) * chunk_stride + (q % chunk_rows) * out_stride
  long long out_stride;
  int chunk_rows;
  long long chunk_stride;
};
}  // namespace

static int dense_topk_impl(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, int k,
                           long long id_offset, float* out_scores, int32_t* out_ids, const OutLayout& ol,
                           int32_t* out_status, void* ws, size_t ws_bytes, void* stream) {
  QF_REQUIRE(Q >= 0 && N >= 1 && k >= 1, QF_E_INVALID, "qf_dense_topk: bad sizes Q=%d N=%d k=%d", Q, N, k);
  QF_REQUIRE(D >= BK && D % BK == 0 && D <= MAX_D, QF_E_UNSUPPORTED,
             "qf_dense_topk: D=%d must be a multiple of %d and <= %d (query tile lives in tensor memory)", D, BK, MAX_D);
  QF_REQUIRE(k <= qf_dense_max_k(), QF_E_UNSUPPORTED, "qf_dense_topk: k=%d exceeds %d", k, qf_dense_max_k());
  if (Q == 0) return QF_OK;
  QF_REQUIRE(queries_bf16 && corpus_bf16 && out_scores && out_ids && ws, QF_E_INVALID, "qf_dense_topk: null pointer argument");
  QF_REQUIRE((reinterpret_cast<uintptr_t>(corpus_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(queries_bf16) & 15) == 0,
             QF_E_INVALID, "qf_dense_topk: queries/corpus must be 16-byte aligned");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const DenseWs w = plan_ws(Q, N, k);
  QF_REQUIRE(w.n_chunks * 2 < SEL_CAP, QF_E_UNSUPPORTED, "qf_dense_topk: corpus too large for one launch (%d chunks)", w.n_chunks);
  QF_REQUIRE(ws_bytes >= w.total, QF_E_WORKSPACE, "qf_dense_topk: workspace %zu < %zu bytes", ws_bytes, w.total);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned char* base = reinterpret_cast<unsigned char*>(ws);

  DenseMaps maps;
  rc = make_corpus_tmap(corpus_bf16, N, D, BN, &maps.one);
  if (rc != QF_OK) return rc;
  rc = make_corpus_tmap(corpus_bf16, N, D, BN / 2, &maps.pair);
  if (rc != QF_OK) return rc;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[qf::device_index()];   // kernel attributes are per device
  if (!configured) {
    QF_CUDA_TRY(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL_SMEM));
    configured = true;
  }
  const int sms = qf::sm_count();
  const int n_mtiles = (Q + BM - 1) / BM;
  const int T = (int)(((long long)N + BN - 1) / BN);
  float* tau = reinterpret_cast<float*>(base + w.off_tau);

  DenseParams p{};
  p.queries = reinterpret_cast<const __nv_bfloat16*>(queries_bf16);
  p.Q = Q; p.N = N; p.D = D; p.n_mtiles = n_mtiles;
  p.tau = tau;
  p.seg_pairs = reinterpret_cast<uint2*>(base + w.off_segpairs);
  p.seg_cnt = reinterpret_cast<int*>(base + w.off_segcnt);
  p.q_pad = n_mtiles * BM;
  p.n_segs = w.n_chunks * 2;
  p.store = reinterpret_cast<float*>(base + w.off_sample);
  p.store_ld = w.sample_ld;

  auto grid_for = [&](int n_units) { return n_units < sms ? n_units : sms; };
  if (w.n_sample_tiles > 0) {
    // threshold estimation: score a strided sample of doc tiles, take an exact order statistic per query
    const int target = cand_target(k);
    p.n_vtiles = w.n_sample_tiles;
    p.tile_step = T / w.n_sample_tiles;
    p.chunk_tiles = 16;
    // valid sample documents and the order statistic that should admit ~target candidates
    long long s_valid = 0;
    for (int t = 0; t < p.n_vtiles; ++t) {
      const long long d0 = (long long)t * p.tile_step * BN;
      s_valid += std::max(0ll, std::min((long long)BN, (long long)N - d0));
    }
    long long jj = ((long long)target * s_valid + N - 1) / N;
    if (jj < 1) jj = 1;
    if (jj > s_valid) jj = s_valid;
    // groups of 32 documents (one maximum per epilogue thread and tile) when the sample is large; small corpora --
    // e.g. one of 8 shards of the NQ corpus, 335 k rows -- get 16- or 8-document groups so that the cheap estimator
    // still has >= 4 jj values to choose from (round 1 fell back to the exact select over stored scores there: the
    // "tau" phase took 0.18 ms instead of 0.07 ms at N = 8)
    int parts = 1;
    while (parts < 4 && (long long)p.n_vtiles * 2 * parts < 4 * jj) parts *= 2;
    const long long n_vals = (long long)p.n_vtiles * 2 * parts;
    const bool tilemax = jj <= 64 && n_vals <= 1024 && n_vals >= 4 * jj;
    p.mode = tilemax ? MODE_TILEMAX : MODE_STORE;
    p.tm_parts = parts;
    if (tilemax) p.store_ld = n_vals;
    const int units = ((p.n_vtiles + p.chunk_tiles - 1) / p.chunk_tiles) * n_mtiles;
    qf::prof_mark(QF_PROF_DENSE_SAMPLE, 0, st);
    rc = launch_dense(maps, p, st);
    qf::prof_mark(QF_PROF_DENSE_SAMPLE, 1, st);
    if (rc != QF_OK) return rc;
    const long long s_cols = (long long)p.n_vtiles * BN;
    const bool small_j = ((long long)target * s_cols + N - 1) / N <= 128;
    qf::prof_mark(QF_PROF_DENSE_TAU, 0, st);
    if (tilemax) {
      tau_kernel_tilemax<<<Q, 512, 0, st>>>(p.store, p.store_ld, (int)n_vals, (int)jj, Q, tau);
    } else {
      if (small_j) tau_kernel_small<<<Q, 512, 0, st>>>(p.store, p.store_ld, p.n_vtiles, p.tile_step, N, Q, target, tau);
      tau_kernel<<<Q, 512, 0, st>>>(p.store, p.store_ld, p.n_vtiles, p.tile_step, N, Q, target, tau, small_j ? 1 : 0);
    }
    qf::prof_mark(QF_PROF_DENSE_TAU, 1, st);
    p.store_ld = w.sample_ld;
    QF_CUDA_TRY(cudaGetLastError());
  } else {
    fill_kernel<<<(Q + 255) / 256, 256, 0, st>>>(tau, Q, -INFINITY);
    QF_CUDA_TRY(cudaGetLastError());
  }
  p.mode = MODE_FILTER;
  p.n_vtiles = T;
  p.tile_step = 1;
  p.chunk_tiles = w.chunk_tiles;
  {
    const int units = ((p.n_vtiles + p.chunk_tiles - 1) / p.chunk_tiles) * n_mtiles;
    qf::prof_mark(QF_PROF_DENSE_MAIN, 0, st);
    rc = launch_dense(maps, p, st);
    qf::prof_mark(QF_PROF_DENSE_MAIN, 1, st);
    if (rc != QF_OK) return rc;
  }
  SelectParams sp{};
  sp.seg_pairs = p.seg_pairs; sp.seg_cnt = p.seg_cnt; sp.n_segs = w.n_chunks * 2;
  sp.Q = Q; sp.k = k; sp.N_total = N; sp.id_offset = id_offset;
  sp.out_scores = out_scores; sp.out_ids = out_ids; sp.out_status = out_status;
  sp.out_stride = ol.out_stride; sp.chunk_rows = ol.chunk_rows; sp.chunk_stride = ol.chunk_stride;
  qf::prof_mark(QF_PROF_DENSE_SELECT, 0, st);
  {
    // common case: select2_kernel (3 queries in flight per SM); select_kernel then finishes what it deferred
    static bool s2_dev[64] = {};
    bool& s2_ready = s2_dev[qf::device_index()];
    if (!s2_ready) {
      QF_CUDA_TRY(cudaFuncSetAttribute(select2_kernel<S2_CAP, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)S2Layout<S2_CAP>::total));
      QF_CUDA_TRY(cudaFuncSetAttribute(select2_kernel<S2_CAP, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      s2_ready = true;
    }
    static const bool no_s2 = getenv("QF_DENSE_NO_SELECT2") != nullptr;     // debug hook: old ranking kernel only
    if (!no_s2 && w.n_chunks * 2 <= S2_MAXSEG && k >= 1) {
      const int g2 = Q < 3 * sms ? Q : 3 * sms;
      select2_kernel<S2_CAP, 3><<<g2, S2_NT, S2Layout<S2_CAP>::total, st>>>(sp);
      QF_CUDA_TRY(cudaGetLastError());
      sp.fixup = 1;
    }
  }
  select_kernel<<<Q < sms ? Q : sms, SEL_NT, SEL_SMEM, st>>>(sp);
  qf::prof_mark(QF_PROF_DENSE_SELECT, 1, st);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

// Full score matrix for a (small) query batch: out[q][d] = <queries[q], corpus[d]>, rows padded to a multiple of
// 128 and columns to a multiple of 64 (ld).  Same tcgen05 kernel, STORE epilogue.
extern "C" int qf_dense_scores(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, float* out,
                               long long ld, void* stream) {
  QF_REQUIRE(Q >= 1 && N >= 1, QF_E_INVALID, "qf_dense_scores: bad sizes Q=%d N=%d", Q, N);
  QF_REQUIRE(D >= BK && D % BK == 0 && D <= MAX_D, QF_E_UNSUPPORTED, "qf_dense_scores: D=%d unsupported", D);
  const long long T = ((long long)N + BN - 1) / BN;
  QF_REQUIRE(ld >= T * BN && ld % 4 == 0, QF_E_INVALID, "qf_dense_scores: ld=%lld must be >= %lld and a multiple of 4", ld, T * BN);
  QF_REQUIRE(queries_bf16 && corpus_bf16 && out, QF_E_INVALID, "qf_dense_scores: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  DenseMaps maps;
  rc = make_corpus_tmap(corpus_bf16, N, D, BN, &maps.one);
  if (rc != QF_OK) return rc;
  rc = make_corpus_tmap(corpus_bf16, N, D, BN / 2, &maps.pair);
  if (rc != QF_OK) return rc;
  DenseParams p{};
  p.queries = reinterpret_cast<const __nv_bfloat16*>(queries_bf16);
  p.Q = Q; p.N = N; p.D = D;
  p.n_mtiles = (Q + BM - 1) / BM;
  p.mode = MODE_STORE;
  p.n_vtiles = (int)T;
  p.tile_step = 1;
  p.chunk_tiles = T >= 4096 ? 256 : (T >= 256 ? 32 : 4);
  p.store = out;
  p.store_ld = ld;
  return launch_dense(maps, p, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int qf_dense_row_topk(const float* scores, long long ld, const int32_t* rows, int n_rows, int N, int k,
                                 long long id_offset, float* out_scores, int32_t* out_ids, int32_t* out_status,
                                 void* stream) {
  QF_REQUIRE(n_rows >= 0 && N >= 1 && k >= 1 && k <= qf_dense_max_k(), QF_E_INVALID, "qf_dense_row_topk: bad sizes");
  if (n_rows == 0) return QF_OK;
  QF_REQUIRE(scores && rows && out_scores && out_ids && ld >= N, QF_E_INVALID, "qf_dense_row_topk: bad argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  QF_CUDA_TRY(cudaFuncSetAttribute(row_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL_SMEM));
  RowTopkParams rp{scores, ld, rows, n_rows, N, k, id_offset, out_scores, out_ids, out_status};
  const int sms = qf::sm_count();
  row_topk_kernel<<<n_rows < sms ? n_rows : sms, SEL_NT, SEL_SMEM, reinterpret_cast<cudaStream_t>(stream)>>>(rp);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

extern "C" int qf_dense_topk(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, int k,
                             long long id_offset, float* out_scores, int32_t* out_ids, int32_t* out_status, void* ws,
                             size_t ws_bytes, void* stream) {
  return dense_topk_impl(queries_bf16, corpus_bf16, Q, N, D, k, id_offset, out_scores, out_ids, OutLayout{k, 0, 0},
                         out_status, ws, ws_bytes, stream);
}

extern "C" int qf_dense_topk_packed(const void* queries_bf16, const void* corpus_bf16, int Q, int N, int D, int k,
                                    long long id_offset, int32_t* out_pairs, int chunk_rows, long long chunk_stride,
                                    int32_t* out_status, void* ws, size_t ws_bytes, void* stream) {
  QF_REQUIRE(out_pairs != nullptr && chunk_rows >= 0 && (chunk_rows == 0 || chunk_stride >= 2LL * k * chunk_rows),
             QF_E_INVALID, "qf_dense_topk_packed: bad output layout (chunk_rows=%d chunk_stride=%lld)", chunk_rows,
             chunk_stride);
  return dense_topk_impl(queries_bf16, corpus_bf16, Q, N, D, k, id_offset, reinterpret_cast<float*>(out_pairs),
                         out_pairs + k, OutLayout{2LL * k, chunk_rows, chunk_stride}, out_status, ws, ws_bytes, stream);
}

static int topk_merge_impl(const float* scores, const int32_t* ids, int G, int Q, int k, long long q_stride,
                           long long seg_stride, float* out_scores, int32_t* out_ids, long long out_stride, int k_out,
                           void* stream);

extern "C" int qf_topk_merge(const float* scores, const int32_t* ids, int G, int Q, int k, float* out_scores,
                             int32_t* out_ids, void* stream) {
  return topk_merge_impl(scores, ids, G, Q, k, k, (long long)Q * k, out_scores, out_ids, k, k, stream);
}

extern "C" int qf_topk_merge_packed(const int32_t* in_pairs, int G, int Q, int k_in, long long seg_stride, int k_out,
                                    int32_t* out_pairs, void* stream) {
  QF_REQUIRE(in_pairs != nullptr && out_pairs != nullptr && seg_stride >= 2LL * k_in * Q && k_out >= 1, QF_E_INVALID,
             "qf_topk_merge_packed: bad argument (seg_stride=%lld k_out=%d)", seg_stride, k_out);
  return topk_merge_impl(reinterpret_cast<const float*>(in_pairs), in_pairs + k_in, G, Q, k_in, 2LL * k_in, seg_stride,
                         reinterpret_cast<float*>(out_pairs), out_pairs + k_out, 2LL * k_out, k_out, stream);
}

// Completeness check of a merge over TRUNCATED shard lists (k_in < k_out entries per shard): the merged top k_out is
// exact unless some shard's list is full and its last score reaches the merged k_out-th score (that shard may hold
// further documents at or above it), or the merged list itself is not full while a shard's list is.
__global__ void merge_check_kernel(const int32_t* in_pairs, int G, int Q, int k_in, long long seg_stride,
                                   const int32_t* merged, int k_out, int32_t* flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)G * Q) return;
  const int g = (int)(i / Q), q = (int)(i % Q);
  const int32_t* row = in_pairs + (size_t)g * seg_stride + (size_t)q * 2 * k_in;
  const int last_id = row[k_in + k_in - 1];
  if (last_id < 0) return;                                   // the shard sent everything it has for this query
  const float last = __int_as_float(row[k_in - 1]);
  const int32_t* mr = merged + (size_t)q * 2 * k_out;
  const bool full = mr[k_out + k_out - 1] >= 0;
  if (!full || last >= __int_as_float(mr[k_out - 1])) atomicOr(flag, 1);
}

extern "C" int qf_topk_merge_check(const int32_t* in_pairs, int G, int Q, int k_in, long long seg_stride,
                                   const int32_t* merged_pairs, int k_out, int32_t* flag, void* stream) {
  QF_REQUIRE(in_pairs && merged_pairs && flag && G >= 1 && Q >= 0 && k_in >= 1 && k_out >= 1, QF_E_INVALID,
             "qf_topk_merge_check: bad argument");
  if (Q == 0) return QF_OK;
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  const long long n = (long long)G * Q;
  merge_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      in_pairs, G, Q, k_in, seg_stride, merged_pairs, k_out, flag);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

static int topk_merge_impl(const float* scores, const int32_t* ids, int G, int Q, int k, long long q_stride,
                           long long seg_stride, float* out_scores, int32_t* out_ids, long long out_stride, int k_out,
                           void* stream) {
  QF_REQUIRE(G >= 1 && Q >= 0 && k >= 1, QF_E_INVALID, "qf_topk_merge: bad sizes G=%d Q=%d k=%d", G, Q, k);
  QF_REQUIRE((long long)G * k <= SEL_CAP, QF_E_UNSUPPORTED, "qf_topk_merge: G*k=%lld exceeds %d", (long long)G * k, SEL_CAP);
  QF_REQUIRE(k_out <= qf_dense_max_k(), QF_E_UNSUPPORTED, "qf_topk_merge: k_out=%d exceeds %d", k_out, qf_dense_max_k());
  if (Q == 0) return QF_OK;
  QF_REQUIRE(scores && ids && out_scores && out_ids, QF_E_INVALID, "qf_topk_merge: null pointer argument");
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[qf::device_index()];   // kernel attributes are per device
  if (!configured) {
    QF_CUDA_TRY(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL_SMEM));
    QF_CUDA_TRY(cudaFuncSetAttribute(select2_kernel<S2_CAP_MERGE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)S2Layout<S2_CAP_MERGE>::total));
    QF_CUDA_TRY(cudaFuncSetAttribute(select2_kernel<S2_CAP_MERGE, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  SelectParams sp{};
  sp.seg_pairs = nullptr; sp.seg_cnt = nullptr; sp.cand_score = scores; sp.cand_id = ids;
  sp.q_stride = q_stride; sp.seg_stride = seg_stride; sp.nseg = G; sp.seg_len = k;
  sp.Q = Q; sp.k = k_out; sp.N_total = 0x7fffffff; sp.id_offset = 0;
  sp.out_scores = out_scores; sp.out_ids = out_ids; sp.out_status = nullptr;
  sp.out_stride = out_stride; sp.chunk_rows = 0; sp.chunk_stride = 0;
  const int sms = qf::sm_count();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const bool no_s2 = getenv("QF_DENSE_NO_SELECT2") != nullptr;     // debug hook: old ranking kernel only
  if (!no_s2) {
    // common case: bucket-cut kernel, two queries in flight per SM; select_kernel then finishes what it deferred
    const int g2 = Q < 2 * sms ? Q : 2 * sms;
    select2_kernel<S2_CAP_MERGE, 2><<<g2, S2_NT, S2Layout<S2_CAP_MERGE>::total, st>>>(sp);
    QF_CUDA_TRY(cudaGetLastError());
    sp.fixup = 1;
  }
  select_kernel<<<Q < sms ? Q : sms, SEL_NT, SEL_SMEM, st>>>(sp);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}
