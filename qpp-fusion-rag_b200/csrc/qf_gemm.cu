// qf_gemm: the Linear layers of the query encoder on the 5th-generation tensor cores, own kernel (round 2; round 1 called
// cuBLASLt here).
//
//   Y[M, N] (bf16) = act( X[M, K] (bf16, row-major) . W[N, K]^T (bf16, row-major = nn.Linear weight) + bias[N] (bf16) )
//   act = identity | erf-GELU (fused: the FFN-1 activation no longer needs its own pass over [M, 4H])
//
// replaces the four nn.Linear forward calls per BERT layer inside SentenceTransformer.encode
// (reference src/retrievers/bge.py:153-170 -> transformers BertLayer).
//
// One persistent CTA per SM, 6 warps, output tiles of 128 x BN (BN = 256 or 128):
//   warp 0   TMA producer: X and W tiles (64 k-elements = 128-byte rows, SWIZZLE_128B) into a 4-stage ring
//            (cp.async.bulk.tensor.2d + mbarrier expect_tx)
//   warp 1   MMA issuer: tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN, K = 16), both operands from shared-memory
//            descriptors, fp32 accumulators in TENSOR MEMORY -- two accumulator stages of BN columns, so the epilogue of
//            tile i overlaps the MMAs of tile i + 1; tcgen05.commit frees ring slots / publishes accumulator stages
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction), + bias, optional GELU, bf16 -> swizzled
//            staging buffer -> TMA store (cp.async.bulk.tensor, [128 x 64] boxes), double-buffered
// Tile order: n fastest, so the CTAs running at any moment share a few X row blocks and all of W (L2-resident).
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>
#include <vector>

#include "qf_common.cuh"
#include "qf_gemm.cuh"

namespace {

using u32 = unsigned int;
using u64 = unsigned long long;
using bf16 = __nv_bfloat16;

constexpr int GM = 128;          // rows per tile (TMEM lanes)
constexpr int GK = 64;           // k elements per stage (128-byte rows)
constexpr int GSTAGES = 4;
constexpr int GNT = 192;         // 6 warps

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
// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16 = bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_ss(u32 tmem_d, u64 desc_a, u64 desc_b, u32 idesc, u32 accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(u32 bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
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
// TMA store of one [128 rows x 64 columns] bf16 box (SWIZZLE_128B staging) + bulk-group bookkeeping
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, u32 smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<u64>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_PENDING) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps
__device__ __forceinline__ void sts128(u32 sa, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ u32 elect_one() {
  u32 pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// float32 erf-GELU (same rational approximation as qf_bert.cu's gelu_kernel: |GELU error| <= 1.4e-6)
__device__ __forceinline__ float gelu_erf(float x) {
  float z = x * 0.70710678118654752f;
  z = fminf(fmaxf(z, -4.0f), 4.0f);
  const float z2 = z * z;
  float p = fmaf(-2.72614225801306e-10f, z2, 2.77068142495902e-08f);
  p = fmaf(p, z2, -2.10102402082508e-06f);
  p = fmaf(p, z2, -5.69250639462346e-05f);
  p = fmaf(p, z2, -7.34990630326855e-04f);
  p = fmaf(p, z2, -2.95459980854025e-03f);
  p = fmaf(p, z2, -1.60960333262415e-02f);
  float q = fmaf(-1.45660718464996e-05f, z2, -2.13374055278905e-04f);
  q = fmaf(q, z2, -1.68282697438203e-03f);
  q = fmaf(q, z2, -7.37332916720468e-03f);
  q = fmaf(q, z2, -1.42647390514189e-02f);
  const float e = __fdividef(z * p, q);
  const float h = 0.5f * x;
  return fmaf(h, e, h);
}

struct __align__(8) GemmCtl {
  u64 full[GSTAGES];
  u64 empty[GSTAGES];
  u64 tmem_full[2];
  u64 tmem_empty[2];
  u32 tmem_base;
  u32 pad;
};

struct GemmParams {
  const bf16* bias;     // [N]
  bf16* Y;              // [M, N]
  int M, N, K;
  int act;              // 0 none, 1 erf-GELU
};

constexpr int EPI_COLS = 64;                          // columns per staged store box (128-byte rows)
constexpr int EPI_BYTES = GM * EPI_COLS * 2;          // 16 KB per staging buffer, two of them

template <int BN>
constexpr size_t gemm_smem() {
  return (size_t)GSTAGES * (GM * GK * 2 + BN * GK * 2) + 2 * EPI_BYTES + sizeof(GemmCtl) + 1024;
}

template <int BN>
__global__ void __launch_bounds__(GNT, 1) gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB,
                                                      const __grid_constant__ CUtensorMap tmY, const GemmParams p) {
  constexpr int A_BYTES = GM * GK * 2, B_BYTES = BN * GK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128
  constexpr u32 IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(BN >> 3) << 17) | ((u32)(GM >> 4) << 24);
  constexpr u32 TMEM_COLS = 2 * BN;                 // 512 or 256: two accumulator stages
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  GemmCtl* ctl = reinterpret_cast<GemmCtl*>(smem + GSTAGES * STAGE_BYTES + 2 * EPI_BYTES);
  const u32 a_epi = smem_u32(smem + GSTAGES * STAGE_BYTES);           // two 16 KB store staging buffers (1024-aligned)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles_n = p.N / BN, n_tiles = (p.M / GM) * n_tiles_n, n_kb = p.K / GK;
  const u32 a_stage0 = smem_u32(smem);
  const u32 a_full = smem_u32(&ctl->full[0]), a_empty = smem_u32(&ctl->empty[0]);
  const u32 a_tfull = smem_u32(&ctl->tmem_full[0]), a_tempty = smem_u32(&ctl->tmem_empty[0]);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < GSTAGES; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->tmem_full[i], 1);
      mbar_init(&ctl->tmem_empty[i], 4);            // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const u32 tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // (the whole warp walks the loop, one elected lane issues: the warp stays converged for the final barrier)
    u32 slot = 0, phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (t / n_tiles_n) * GM, n0 = (t % n_tiles_n) * BN;
      for (int kb = 0; kb < n_kb; ++kb) {
        if (elect_one()) {
          mbar_wait(a_empty + 8 * slot, phase ^ 1);                  // slot free (first pass: immediately)
          mbar_expect_tx(a_full + 8 * slot, STAGE_BYTES);
          const u32 dst = a_stage0 + slot * STAGE_BYTES;
          tma_load_2d(dst, &tmA, a_full + 8 * slot, kb * GK, m0);
          tma_load_2d(dst + A_BYTES, &tmB, a_full + 8 * slot, kb * GK, n0);
        }
        __syncwarp();
        if (++slot == GSTAGES) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    u32 slot = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(a_tempty + 8 * acc, acc_phase ^ 1);                  // epilogue has drained this accumulator stage
      tc_fence_after();
      const u32 tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(a_full + 8 * slot, phase);
        tc_fence_after();
        if (elect_one()) {
          const u64 da = make_desc_sw128(a_stage0 + slot * STAGE_BYTES);
          const u64 db = make_desc_sw128(a_stage0 + slot * STAGE_BYTES + A_BYTES);
#pragma unroll
          for (int k = 0; k < GK / 16; ++k)
            umma_ss(tmem_d, da + (u64)(k * 2), db + (u64)(k * 2), IDESC, (kb | k) != 0);
          umma_commit(a_empty + 8 * slot);                            // slot reusable once these MMAs have read it
        }
        __syncwarp();
        if (++slot == GSTAGES) { slot = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(a_tfull + 8 * acc);                // accumulator stage complete
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    // 64 columns at a time: tensor memory -> registers -> + bias [-> GELU] -> bf16 -> a SWIZZLE_128B staging buffer
    // (each thread writes its row's 128 bytes as eight 16-byte chunks at chunk ^ (row & 7): conflict-free) -> ONE TMA
    // store of the [128 x 64] box by an elected thread.  Two staging buffers: the store of chunk c overlaps the math of
    // chunk c + 1; a buffer is rewritten only after its store has READ it (cp.async.bulk.wait_group.read).
    const u32 lane_base = (u32)((warp & 3) * 32) << 16;               // the TMEM lane quarter this warp may access
    const int row_in_tile = (warp & 3) * 32 + lane;
    const bool issuer = (warp == 2 && lane == 0);
    u32 acc = 0, acc_phase = 0, buf = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (t / n_tiles_n) * GM, n0 = (t % n_tiles_n) * BN;
      mbar_wait(a_tfull + 8 * acc, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / EPI_COLS; ++c) {
        u32 v0[32], v1[32];
        tmem_ld32(tmem_base + lane_base + acc * BN + c * EPI_COLS, v0);
        tmem_ld32(tmem_base + lane_base + acc * BN + c * EPI_COLS + 32, v1);
        if (issuer) bulk_wait_read<1>();                               // the store that last used this buffer has read it
        epi_barrier();
        tmem_wait_ld();
        const uint4* b4 = reinterpret_cast<const uint4*>(p.bias + n0 + c * EPI_COLS);
        const u32 srow = a_epi + buf * EPI_BYTES + (u32)row_in_tile * 128u;
#pragma unroll
        for (int g = 0; g < 8; ++g) {                                  // 8 columns = one 16-byte chunk
          const uint4 bb = __ldg(b4 + g);
          const u32 bw[4] = {bb.x, bb.y, bb.z, bb.w};
          u32 out[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int col = g * 8 + 2 * h;
            const float b0 = __uint_as_float(bw[h] << 16), b1 = __uint_as_float(bw[h] & 0xFFFF0000u);
            float x0 = __uint_as_float(col < 32 ? v0[col & 31] : v1[col & 31]) + b0;
            float x1 = __uint_as_float(col < 32 ? v0[(col + 1) & 31] : v1[(col + 1) & 31]) + b1;
            if (p.act == 1) {
              x0 = gelu_erf(x0);
              x1 = gelu_erf(x1);
            }
            const __nv_bfloat162 pk = __floats2bfloat162_rn(x0, x1);
            out[h] = *reinterpret_cast<const u32*>(&pk);
          }
          sts128(srow + (u32)((g ^ (row_in_tile & 7)) * 16), make_uint4(out[0], out[1], out[2], out[3]));
        }
        fence_async_smem();                                            // generic-proxy writes -> visible to the TMA unit
        epi_barrier();
        if (issuer) {
          tma_store_2d(&tmY, a_epi + buf * EPI_BYTES, n0 + c * EPI_COLS, m0);
          bulk_commit();
        }
        buf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (issuer) bulk_wait_all();                                       // all stores complete before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeTiledFn* out) {
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

// rows x cols bf16 row-major matrix, boxes of box_rows x 64 columns, SWIZZLE_128B.  Encoded maps are cached (a tensor
// map is a pure function of these five values): an encoder call issues 48 GEMMs = 144 maps, all of which repeat from
// call to call as long as the allocator hands back the same blocks.
struct TmapKey {
  const void* base;
  long long rows;
  int cols, box_rows, dev;
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap map;
};
std::mutex g_tmap_mu;
std::vector<TmapEntry> g_tmaps;

int make_tmap(const void* base, long long rows, int cols, int box_rows, CUtensorMap* tmap) {
  const int dev = qf::device_index();
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    for (const TmapEntry& e : g_tmaps)
      if (e.key.base == base && e.key.rows == rows && e.key.cols == cols && e.key.box_rows == box_rows && e.key.dev == dev) {
        *tmap = e.map;
        return QF_OK;
      }
  }
  EncodeTiledFn encode;
  int rc = get_encode(&encode);
  if (rc != QF_OK) return rc;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)GK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  QF_REQUIRE(r == CUDA_SUCCESS, QF_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmaps.size() >= 1024) g_tmaps.erase(g_tmaps.begin(), g_tmaps.begin() + 512);
    g_tmaps.push_back(TmapEntry{TmapKey{base, rows, cols, box_rows, dev}, *tmap});
  }
  return QF_OK;
}

template <int BN>
int launch_gemm(const bf16* W, const bf16* X, const GemmParams& p, cudaStream_t st) {
  CUtensorMap tmA, tmB, tmY;
  int rc = make_tmap(X, p.M, p.K, GM, &tmA);
  if (rc != QF_OK) return rc;
  rc = make_tmap(W, p.N, p.K, BN, &tmB);
  if (rc != QF_OK) return rc;
  rc = make_tmap(p.Y, p.M, p.N, GM, &tmY);            // store boxes: 128 rows x 64 columns
  if (rc != QF_OK) return rc;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[qf::device_index()];
  if (!configured) {
    QF_CUDA_TRY(cudaFuncSetAttribute(gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<BN>()));
    configured = true;
  }
  const int n_tiles = (p.M / GM) * (p.N / BN);
  const int sms = qf::sm_count();
  gemm_kernel<BN><<<n_tiles < sms ? n_tiles : sms, GNT, gemm_smem<BN>(), st>>>(tmA, tmB, tmY, p);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

}  // namespace

namespace qf {

bool gemm_supported(long long M, int N, int K) {
  return M > 0 && M % GM == 0 && M <= 0x7fffffffLL && N % 128 == 0 && K % GK == 0 && K >= GK;
}

int gemm_bias_act(const __nv_bfloat16* W, const __nv_bfloat16* X, const __nv_bfloat16* bias, __nv_bfloat16* Y, long long M,
                  int N, int K, int act, cudaStream_t st) {
  QF_REQUIRE(gemm_supported(M, N, K), QF_E_UNSUPPORTED, "qf_gemm: M=%lld N=%d K=%d outside the envelope (M %% 128, N %% 128, K %% 64)",
             M, N, K);
  QF_REQUIRE(((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) |
               reinterpret_cast<uintptr_t>(bias)) & 15) == 0, QF_E_INVALID, "qf_gemm: operands must be 16-byte aligned");
  GemmParams p{bias, Y, (int)M, N, K, act};
  return N % 256 == 0 ? launch_gemm<256>(W, X, p, st) : launch_gemm<128>(W, X, p, st);
}

}  // namespace qf

// C ABI: one Linear layer (testing / other hosts); the encoder calls qf::gemm_bias_act directly.
extern "C" int qf_linear_bf16(const void* W, const void* X, const void* bias, void* Y, long long M, int N, int K, int act,
                              void* stream) {
  QF_REQUIRE(W && X && bias && Y, QF_E_INVALID, "qf_linear_bf16: null pointer argument");
  QF_REQUIRE(act == 0 || act == 1, QF_E_INVALID, "qf_linear_bf16: unknown activation %d", act);
  int rc = qf::check_device();
  if (rc != QF_OK) return rc;
  return qf::gemm_bias_act(reinterpret_cast<const __nv_bfloat16*>(W), reinterpret_cast<const __nv_bfloat16*>(X),
                           reinterpret_cast<const __nv_bfloat16*>(bias), reinterpret_cast<__nv_bfloat16*>(Y), M, N, K, act,
                           reinterpret_cast<cudaStream_t>(stream));
}
