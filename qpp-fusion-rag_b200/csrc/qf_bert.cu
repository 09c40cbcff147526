// qf_bert: BERT-base query encoder on the device -- the step right before the dense search
// (reference: BGERetriever._encode_queries -> SentenceTransformer("BAAI/bge-base-en-v1.5").encode(...,
// normalize_embeddings=True), src/retrievers/bge.py:153-170; model = HF BertModel + CLS pooling + L2 normalisation).
//
// Activations are bf16 [B*L, H] row-major (one row per token, padded to L tokens per query), statistics and
// accumulation fp32.  The four Linear layers of a block run on the library's own tcgen05 GEMM (csrc/qf_gemm.cu: TMA loads,
// TMEM accumulators, bias + erf-GELU in the epilogue; since round 2); cuBLASLt is kept only for shapes outside that
// kernel's envelope (row counts not a multiple of 128 never occur: the token rows are padded).  Between the GEMMs:
//   embed_ln      word + position + type embedding -> LayerNorm                       (warp per token, 16-byte loads)
//   attn          L <= 128: one WARP per (query, head) on mma.sync tiles, K/V staged in the warp's shared-memory slice,
//                 masked softmax on the accumulator fragments, which then ARE the A operand of P.V;
//                 longer sequences: one CTA per (query, head), warp per query row, P.V through shuffles.
//                 Either way the [L, L] score matrix never leaves the SM and padded query rows are not computed
//   add_ln        residual + LayerNorm                                                (warp per token)
//   gelu          erf GELU in place (float32 rational erf; cuBLASLt's fused GELU is the tanh approximation, HF BERT uses erf)
//   pool_norm     CLS (or masked mean) pooling + L2 normalisation -> bf16 queries for qf_dense_topk and/or fp32
// Weights: ONE packed bf16 blob, layout in include/qppfusion.h (qf_bert_weight_elems).
#include <cublasLt.h>
#include <cuda_bf16.h>

#include <mutex>
#include <vector>

#include "qf_common.cuh"
#include "qf_gemm.cuh"

namespace {

using bf16 = __nv_bfloat16;

struct Offsets {
  size_t word, pos, type, eg, eb, layer0, per_layer, total;
  // inside a layer
  size_t wqkv, bqkv, wo, bo, g1, b1, w1, bi1, w2, bi2, g2, b2;
};

Offsets offsets(const qf_bert_config& c) {
  Offsets o{};
  const size_t H = c.hidden, F = c.ffn;
  size_t a = 0;
  o.word = a; a += (size_t)c.vocab * H;
  o.pos = a;  a += (size_t)c.max_pos * H;
  o.type = a; a += 2 * H;
  o.eg = a;   a += H;
  o.eb = a;   a += H;
  o.layer0 = a;
  size_t l = 0;
  o.wqkv = l; l += 3 * H * H;
  o.bqkv = l; l += 3 * H;
  o.wo = l;   l += H * H;
  o.bo = l;   l += H;
  o.g1 = l;   l += H;
  o.b1 = l;   l += H;
  o.w1 = l;   l += F * H;
  o.bi1 = l;  l += F;
  o.w2 = l;   l += H * F;
  o.bi2 = l;  l += H;
  o.g2 = l;   l += H;
  o.b2 = l;   l += H;
  o.per_layer = l;
  o.total = a + (size_t)c.n_layers * l;
  return o;
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}


// Where sequence b lives in the [tokens, H] activation matrix.  Padded layout (cu == null): rows [b*L, b*L + L), the
// first `len` are real.  Packed layout (cu = exclusive prefix sums of the lengths, B + 1 entries): rows
// [cu[b], cu[b+1]), all real; L is then only the longest admissible sequence (sizes the shared memory).
struct Seq {
  size_t row0;
  int len;    // real tokens taking part in attention / pooling (clamped to [0, L])
  int rows;   // rows of the matrix that belong to the sequence
};
__device__ __forceinline__ Seq seq_of(const int32_t* lens, const int32_t* cu, int b, int L) {
  Seq q;
  if (cu) {
    const int s0 = __ldg(cu + b), n = __ldg(cu + b + 1) - s0;
    q.row0 = (size_t)s0;
    q.rows = n < 0 ? 0 : n;
    q.len = q.rows > L ? L : q.rows;
  } else {
    int len = lens ? __ldg(lens + b) : L;
    q.row0 = (size_t)b * L;
    q.rows = L;
    q.len = len < 1 ? 1 : (len > L ? L : len);
  }
  return q;
}

// LayerNorm of one token row held as VPL 16-byte vectors per lane (H = 256 * VPL): HF BertLayerNorm = biased variance,
// y = (x - mean) / sqrt(var + eps) * gamma + beta.
template <int VPL>
__device__ __forceinline__ void ln_store(float (&x)[VPL][8], const bf16* gamma, const bf16* beta, float eps, int H,
                                         bf16* out_row, int lane) {
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[v][i];
  const float mean = wsum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = x[v][i] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(wsum(q) / (float)H + eps);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    float g[8], b[8], y[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + col)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(beta + col)), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = (x[v][i] - mean) * rstd * g[i] + b[i];
    *reinterpret_cast<uint4*>(out_row + col) = pack8(y);
  }
}

template <int VPL>
__global__ void __launch_bounds__(256) embed_ln_kernel(const int32_t* tokens, const bf16* word, const bf16* pos,
                                                       const bf16* type, const bf16* gamma, const bf16* beta, float eps,
                                                       bf16* out, long long M, int L, int V, int32_t* status,
                                                       const int32_t* cu, int B, int max_pos) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  constexpr int H = 256 * VPL;
  int b, p;
  if (cu) {                                        // packed: sequence = last b with cu[b] <= m, position inside it
    int lo = 0, hi = B - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if ((long long)__ldg(cu + mid) <= m) lo = mid; else hi = mid - 1;
    }
    b = lo;
    p = (int)(m - __ldg(cu + b));
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
  } else {
    b = (int)(m / L);
    p = (int)(m % L);
  }
  int tok = __ldg(tokens + m);
  if (tok < 0 || tok >= V) {                       // never read outside the table; reported, not repaired silently
    if (lane == 0 && status) atomicOr(status + b, 1);
    tok = 0;
  }
  float x[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    float a[8], b[8], c[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(word + (size_t)tok * H + col)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(pos + (size_t)p * H + col)), b);
    unpack8(__ldg(reinterpret_cast<const uint4*>(type + col)), c);      // token_type_ids = 0
#pragma unroll
    for (int i = 0; i < 8; ++i) x[v][i] = (a[i] + c[i]) + b[i];         // HF: inputs_embeds + token_type, then + position
  }
  ln_store<VPL>(x, gamma, beta, eps, H, out + (size_t)m * H, lane);
}

template <int VPL>
__global__ void __launch_bounds__(256) add_ln_kernel(const bf16* x_in, const bf16* res, const bf16* gamma,
                                                     const bf16* beta, float eps, bf16* out, long long M) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  constexpr int H = 256 * VPL;
  float x[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    float a[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(x_in + (size_t)m * H + col), a);
    unpack8(*reinterpret_cast<const uint4*>(res + (size_t)m * H + col), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[v][i] = a[i] + b[i];
  }
  ln_store<VPL>(x, gamma, beta, eps, H, out + (size_t)m * H, lane);
}

// GELU(x) = x/2 * (1 + erf(x / sqrt 2)) with erf as the float32 rational approximation z * P(z^2) / Q(z^2) on the clamped
// argument (the coefficients Eigen and XLA use for float32 erf; |error| <= 4.5e-7 measured against a float64 erf over
// [-6, 6], GELU error <= 1.4e-6: far below half a bf16 ulp of anything the kernel stores).  erff makes this pass
// instruction-bound (3.9 TB/s); with 13 FMAs and one approximate division it runs at 5.1 TB/s.
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

__global__ void __launch_bounds__(256) gelu_kernel(bf16* x, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    uint4* p = reinterpret_cast<uint4*>(x) + i;
    float f[8];
    unpack8(*p, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = gelu_erf(f[k]);
    *p = pack8(f);
  }
}

// One CTA per (head, query): 4 warps, warp w owns query rows w, w+4, ...; lane l owns keys l, l+32, ... (TL per lane).
constexpr int HD = 64;            // head dimension
constexpr int KPAD = HD + 2;      // K row stride in bf16: 33 words -> lanes reading different rows hit different banks
__host__ __device__ inline size_t attn_q_offset(int L) {
  return (((size_t)L * KPAD + (size_t)L * HD) * sizeof(bf16) + 15) & ~size_t(15);
}
template <int TL>
__global__ void __launch_bounds__(128) attn_kernel(const bf16* qkv, const int32_t* lens, const int32_t* cu, bf16* ctx,
                                                   int L, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* sK = reinterpret_cast<bf16*>(smem_raw);                       // [L][KPAD]
  bf16* sV = sK + (size_t)L * KPAD;                                    // [L][HD]
  float* sQ = reinterpret_cast<float*>(smem_raw + attn_q_offset(L));   // [4][HD], 16-byte aligned for any L
  const int h = blockIdx.y, b = blockIdx.x;                            // sequences on x: no 65535 limit
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Seq sq = seq_of(lens, cu, b, L);
  const int len = sq.len, rows = sq.rows;
  const size_t row0 = sq.row0;
  const int H3 = 3 * H;
  for (int i = tid; i < len * (HD / 2); i += 128) {                    // bf16 pairs of the real keys
    const int j = i / (HD / 2), d2 = i % (HD / 2);
    const bf16* base = qkv + (row0 + j) * H3 + h * HD + 2 * d2;
    *reinterpret_cast<__nv_bfloat162*>(sK + (size_t)j * KPAD + 2 * d2) = *reinterpret_cast<const __nv_bfloat162*>(base + H);
    *reinterpret_cast<__nv_bfloat162*>(sV + (size_t)j * HD + 2 * d2) = *reinterpret_cast<const __nv_bfloat162*>(base + 2 * H);
  }
  __syncthreads();
  float* q = sQ + warp * HD;
  for (int i = warp; i < rows; i += 4) {
    if (i >= len) {                                                    // padding token: never read by a real one
      bf16* out = ctx + (row0 + i) * H + h * HD;
      out[lane] = out[lane + 32] = __float2bfloat16_rn(0.f);
      continue;
    }
    __syncwarp();
    {
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qkv + (row0 + i) * H3 + h * HD + 2 * lane));
      q[2 * lane] = t.x * 0.125f;                                      // 1 / sqrt(64), exact in binary
      q[2 * lane + 1] = t.y * 0.125f;
    }
    __syncwarp();
    float s[TL];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < TL; ++t) {
      const int j = lane + 32 * t;
      float a = -INFINITY;
      if (j < len) {
        a = 0.f;
        const bf16* kr = sK + (size_t)j * KPAD;
#pragma unroll
        for (int d2 = 0; d2 < HD / 2; ++d2) {
          const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(kr + 2 * d2));
          const float2 qv = *reinterpret_cast<const float2*>(q + 2 * d2);
          a = fmaf(qv.x, kv.x, a);
          a = fmaf(qv.y, kv.y, a);
        }
      }
      s[t] = a;
      mx = fmaxf(mx, a);
    }
    mx = wmax(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < TL; ++t) {
      s[t] = (lane + 32 * t < len) ? __expf(s[t] - mx) : 0.f;
      sum += s[t];
    }
    const float inv = 1.0f / wsum(sum);
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int t = 0; t < TL; ++t) {
      const int jn = min(32, len - 32 * t);                            // keys of this slice that are not masked
      for (int jj = 0; jj < jn; ++jj) {
        const float pj = __shfl_sync(0xffffffffu, s[t], jj);
        const bf16* vr = sV + (size_t)(32 * t + jj) * HD;
        o0 = fmaf(pj, __bfloat162float(vr[lane]), o0);
        o1 = fmaf(pj, __bfloat162float(vr[lane + 32]), o1);
      }
    }
    bf16* out = ctx + (row0 + i) * H + h * HD;
    out[lane] = __float2bfloat16_rn(o0 * inv);
    out[lane + 32] = __float2bfloat16_rn(o1 * inv);
  }
}

// Tensor-core attention for query-length sequences (L <= 128): ONE WARP per (query, head), no CTA-wide barrier.
// K and V of the head are staged in the warp's own shared-memory slice (16-byte loads, rows padded to 72 bf16 so both the
// 32-bit B-fragment reads of K and the transposing ldmatrix of V are bank-conflict free); Q fragments come straight from
// global memory.  S = Q.K^T and O = P.V run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate): a 32 x 32 x 64 head is far
// below one tcgen05 tile (M = 128), and the kernel is bound by the QKV read, not by the tensor pipe.  The S accumulators
// of two adjacent key tiles ARE the A fragment of the P.V product, so the probabilities never leave registers.
// Query rows >= len (padding) are not computed: their context rows are written as zeros, and V rows >= len are staged
// as zeros so that 0 x (whatever the padded tokens hold) can never reach a real token.
constexpr int VSTR = HD + 8;      // staged row stride in bf16 (144 bytes)

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <int LT>   // L rounded up to 32, 64 or 128
__global__ void __launch_bounds__(128) attn_mma_kernel(const bf16* __restrict__ qkv, const int32_t* __restrict__ lens,
                                                       const int32_t* __restrict__ cu, bf16* __restrict__ ctx, int L,
                                                       int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = LT / 8;                                            // key tiles of 8
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.y * 4 + warp, b = blockIdx.x;                  // sequences on x: no 65535 limit
  bf16* sK = reinterpret_cast<bf16*>(smem_raw) + (size_t)warp * 2 * LT * VSTR;
  bf16* sV = sK + (size_t)LT * VSTR;
  const Seq sq = seq_of(lens, cu, b, L);
  const int len = sq.len, rows = sq.rows;
  if (rows == 0) return;                                                // (packed layout only) empty sequence
  const size_t row0 = sq.row0;
  const int H3 = 3 * H;
  const int kt_n = (len + 7) >> 3;                                      // key tiles that hold a real key
  const int kt16 = (len + 15) >> 4;                                     // the same in steps of 16 (k steps of P.V)
  {                                                                     // stage K, V rows [0, 16 * kt16)
    const int c8 = lane & 7, r4 = lane >> 3;                            // 8 lanes x 16 bytes = one 64-element row
    for (int j = r4; j < kt16 * 16; j += 4) {
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (j < len) {
        const bf16* base = qkv + (row0 + j) * H3 + h * HD + c8 * 8;
        kv = __ldg(reinterpret_cast<const uint4*>(base + H));
        vv = __ldg(reinterpret_cast<const uint4*>(base + 2 * H));
      }
      *reinterpret_cast<uint4*>(sK + (size_t)j * VSTR + c8 * 8) = kv;
      *reinterpret_cast<uint4*>(sV + (size_t)j * VSTR + c8 * 8) = vv;
    }
  }
  __syncwarp();
  const int g = lane >> 2, t4 = lane & 3;                               // fragment row / column-pair of this lane
  const int mt_n = (len + 15) >> 4;
  for (int mt = 0; mt < mt_n; ++mt) {
    const int ra = mt * 16 + g, rb = ra + 8;                            // the two query rows this lane holds
    const bf16* qa = qkv + (row0 + min(ra, rows - 1)) * H3 + h * HD + t4 * 2;
    const bf16* qb = qkv + (row0 + min(rb, rows - 1)) * H3 + h * HD + t4 * 2;
    uint32_t aq[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      aq[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(qa + ks * 16));
      aq[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(qb + ks * 16));
      aq[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(qa + ks * 16 + 8));
      aq[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(qb + ks * 16 + 8));
    }
    float sc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      if (nt < kt_n) {
        const bf16* kr = sK + (size_t)(nt * 8 + g) * VSTR + t4 * 2;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_bf16(sc[nt], aq[ks], *reinterpret_cast<const uint32_t*>(kr + ks * 16),
                   *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8));
      }
    }
    // masked softmax over the keys; rows ra (elements 0, 1) and rb (elements 2, 3) live in the 4 lanes of a quad
    float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = nt * 8 + t4 * 2 + (e & 1);
        sc[nt][e] = j < len ? sc[nt][e] * 0.125f : -INFINITY;           // 1 / sqrt(64)
      }
      mxa = fmaxf(mxa, fmaxf(sc[nt][0], sc[nt][1]));
      mxb = fmaxf(mxb, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 1));
    mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, 2));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 1));
    mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, 2));
    float sma = 0.f, smb = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      sc[nt][0] = __expf(sc[nt][0] - mxa);                              // exp(-inf) = 0 for the masked keys
      sc[nt][1] = __expf(sc[nt][1] - mxa);
      sc[nt][2] = __expf(sc[nt][2] - mxb);
      sc[nt][3] = __expf(sc[nt][3] - mxb);
      sma += sc[nt][0] + sc[nt][1];
      smb += sc[nt][2] + sc[nt][3];
    }
    sma += __shfl_xor_sync(0xffffffffu, sma, 1);
    sma += __shfl_xor_sync(0xffffffffu, sma, 2);
    smb += __shfl_xor_sync(0xffffffffu, smb, 1);
    smb += __shfl_xor_sync(0xffffffffu, smb, 2);
    const float inva = 1.0f / sma, invb = 1.0f / smb;
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (kk < kt16) {
        uint32_t ap[4];
        ap[0] = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
        ap[1] = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
        ap[2] = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
        ap[3] = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
        // lanes 0-15 address the 16 V rows of this k step; .trans hands every lane V[k pair][d = lane / 4]
        const uint32_t vaddr = (uint32_t)__cvta_generic_to_shared(sV + (size_t)(kk * 16 + (lane & 15)) * VSTR);
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          uint32_t b0, b1;
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
                       : "=r"(b0), "=r"(b1) : "r"(vaddr + dt * 16));
          mma_bf16(o[dt], ap, b0, b1);
        }
      }
    }
    bf16* oa = ctx + (row0 + ra) * H + h * HD + t4 * 2;
    bf16* ob = ctx + (row0 + rb) * H + h * HD + t4 * 2;
    const bool wa = ra < rows, wb = rb < rows;                          // rows in [len, rows) get zeros, beyond: nothing
    const float fa = ra < len ? inva : 0.f, fb = rb < len ? invb : 0.f;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      if (wa) *reinterpret_cast<uint32_t*>(oa + dt * 8) = pack_bf16x2(o[dt][0] * fa, o[dt][1] * fa);
      if (wb) *reinterpret_cast<uint32_t*>(ob + dt * 8) = pack_bf16x2(o[dt][2] * fb, o[dt][3] * fb);
    }
  }
  // context rows of the remaining padding tokens: zeros (8 lanes x 16 bytes per row)
  for (int j = mt_n * 16 + (lane >> 3); j < rows; j += 4)
    *reinterpret_cast<uint4*>(ctx + (row0 + j) * H + h * HD + (lane & 7) * 8) = make_uint4(0, 0, 0, 0);
}

template <int VPL>
__global__ void __launch_bounds__(256) pool_norm_kernel(const bf16* x, const int32_t* lens, const int32_t* cu, int L,
                                                        int pooling, bf16* out_bf16, float* out_f32, int B,
                                                        int32_t* status) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  constexpr int H = 256 * VPL;
  const Seq sq = seq_of(lens, cu, b, L);
  const int len = sq.len;
  if (cu && lane == 0 && status && (sq.rows < 1 || sq.rows > L)) atomicOr(status + b, 2);   // cu_seqlens vs max_len
  float acc[VPL][8];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[v][i] = 0.f;
  const int n_tok = len < 1 ? 0 : (pooling == QF_BERT_POOL_MEAN ? len : 1);    // empty sequence: zero vector
  for (int t = 0; t < n_tok; ++t) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + (sq.row0 + t) * H + (v * 32 + lane) * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[v][i] += f[i];
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[v][i] /= (float)(n_tok > 0 ? n_tok : 1);
      ss += acc[v][i] * acc[v][i];
    }
  const float inv = 1.0f / fmaxf(sqrtf(wsum(ss)), 1e-12f);              // torch.nn.functional.normalize(eps=1e-12)
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = (v * 32 + lane) * 8;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = acc[v][i] * inv;
    if (out_bf16) *reinterpret_cast<uint4*>(out_bf16 + (size_t)b * H + col) = pack8(y);
    if (out_f32) {
      float4* o = reinterpret_cast<float4*>(out_f32 + (size_t)b * H + col);
      o[0] = make_float4(y[0], y[1], y[2], y[3]);
      o[1] = make_float4(y[4], y[5], y[6], y[7]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ cuBLASLt plumbing
constexpr size_t LT_WS = 32u << 20;

struct Plan {
  int dev;
  long long M;
  int N, K;
  cublasLtMatmulDesc_t op;
  cublasLtMatrixLayout_t a, b, c;
  cublasLtMatmulAlgo_t algo;
};
std::mutex g_mu;
cublasLtHandle_t g_lt[64] = {};
std::vector<Plan> g_plans;

#define QF_LT_TRY(expr)                                                                      \
  do {                                                                                       \
    cublasStatus_t _s = (expr);                                                              \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                       \
      qf::set_error("%s failed: cuBLASLt status %d (%s:%d)", #expr, (int)_s, __FILE__, __LINE__); \
      return QF_E_CUDA;                                                                      \
    }                                                                                        \
  } while (0)

// Y[M,N] = X[M,K] . W[N,K]^T + bias[N]   (row-major; as column-major: C[N,M] = op_T(A = W as [K,N]) . B = X as [K,M])
int gemm_bias(const bf16* W, const bf16* X, const bf16* bias, bf16* Y, long long M, int N, int K, void* lt_ws,
              cudaStream_t st) {
  const int dev = qf::device_index();
  Plan plan{};
  bool found = false;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_lt[dev]) QF_LT_TRY(cublasLtCreate(&g_lt[dev]));
    for (const Plan& pl : g_plans)
      if (pl.dev == dev && pl.M == M && pl.N == N && pl.K == K) {
        plan = pl;
        found = true;
        break;
      }
    if (!found) {
      plan.dev = dev; plan.M = M; plan.N = N; plan.K = K;
      QF_LT_TRY(cublasLtMatmulDescCreate(&plan.op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
      const cublasOperation_t T = CUBLAS_OP_T, Nn = CUBLAS_OP_N;
      QF_LT_TRY(cublasLtMatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_TRANSA, &T, sizeof(T)));
      QF_LT_TRY(cublasLtMatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_TRANSB, &Nn, sizeof(Nn)));
      const cublasLtEpilogue_t ep = CUBLASLT_EPILOGUE_BIAS;
      QF_LT_TRY(cublasLtMatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &ep, sizeof(ep)));
      const cudaDataType_t bt = CUDA_R_16BF;
      QF_LT_TRY(cublasLtMatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bt, sizeof(bt)));
      QF_LT_TRY(cublasLtMatrixLayoutCreate(&plan.a, CUDA_R_16BF, K, N, K));
      QF_LT_TRY(cublasLtMatrixLayoutCreate(&plan.b, CUDA_R_16BF, K, M, K));
      QF_LT_TRY(cublasLtMatrixLayoutCreate(&plan.c, CUDA_R_16BF, N, M, N));
      // the heuristic needs the bias pointer attribute to be set to something non-null for the epilogue it plans
      QF_LT_TRY(cublasLtMatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)));
      cublasLtMatmulPreference_t pref;
      QF_LT_TRY(cublasLtMatmulPreferenceCreate(&pref));
      QF_LT_TRY(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &LT_WS, sizeof(LT_WS)));
      cublasLtMatmulHeuristicResult_t res{};
      int n_res = 0;
      const cublasStatus_t hs =
          cublasLtMatmulAlgoGetHeuristic(g_lt[dev], plan.op, plan.a, plan.b, plan.c, plan.c, pref, 1, &res, &n_res);
      cublasLtMatmulPreferenceDestroy(pref);
      if (hs != CUBLAS_STATUS_SUCCESS || n_res == 0) {
        qf::set_error("qf_bert_encode: no cuBLASLt algorithm for M=%lld N=%d K=%d (status %d)", M, N, K, (int)hs);
        return QF_E_UNSUPPORTED;
      }
      plan.algo = res.algo;
      if (g_plans.size() >= 256) {                     // bounded cache: drop the oldest shape
        const Plan& old = g_plans.front();
        cublasLtMatmulDescDestroy(old.op);
        cublasLtMatrixLayoutDestroy(old.a);
        cublasLtMatrixLayoutDestroy(old.b);
        cublasLtMatrixLayoutDestroy(old.c);
        g_plans.erase(g_plans.begin());
      }
      g_plans.push_back(plan);
    }
    // the descriptor carries the bias pointer: set and launch under the lock (launches are asynchronous and short)
    QF_LT_TRY(cublasLtMatmulDescSetAttribute(plan.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)));
    const float one = 1.0f, zero = 0.0f;
    QF_LT_TRY(cublasLtMatmul(g_lt[dev], plan.op, &one, W, plan.a, X, plan.b, &zero, Y, plan.c, Y, plan.c, &plan.algo,
                             lt_ws, LT_WS, st));
  }
  return QF_OK;
}

size_t a256(size_t x) { return (x + 255) & ~size_t(255); }

// Rows the Linear layers run on: M rounded up (128 below 4096 rows, 512 above), so that the packed layout's ever-changing
// token counts map onto few GEMM shapes (one cuBLASLt plan + heuristic query per shape).  The extra rows hold whatever
// the workspace held: rows are independent in every kernel of the encoder and nothing reads them back.
long long gemm_rows(long long M) {
  const long long g = M < 4096 ? 128 : 512;
  return (M + g - 1) / g * g;
}

struct Ws {
  size_t x, qkv, ctx, t1, ff, lt, total;
};
Ws ws_layout(const qf_bert_config& c, long long M_tokens) {
  const long long M = gemm_rows(M_tokens);
  Ws w{};
  size_t a = 0;
  w.x = a;   a += a256((size_t)M * c.hidden * 2);
  w.qkv = a; a += a256((size_t)M * c.hidden * 3 * 2);
  w.ctx = a; a += a256((size_t)M * c.hidden * 2);
  w.t1 = a;  a += a256((size_t)M * c.hidden * 2);
  w.ff = a;  a += a256((size_t)M * c.ffn * 2);
  w.lt = a;  a += LT_WS;
  w.total = a;
  return w;
}

int check_config(const qf_bert_config* c) {
  QF_REQUIRE(c != nullptr, QF_E_INVALID, "qf_bert: null config");
  QF_REQUIRE(c->n_layers >= 1 && c->vocab >= 1 && c->max_pos >= 1 && c->ffn >= 8 && c->ffn % 8 == 0, QF_E_INVALID,
             "qf_bert: bad config (layers=%d vocab=%d max_pos=%d ffn=%d)", c->n_layers, c->vocab, c->max_pos, c->ffn);
  QF_REQUIRE(c->hidden % 256 == 0 && c->hidden >= 256 && c->hidden <= 1024 && c->heads * HD == c->hidden,
             QF_E_UNSUPPORTED, "qf_bert: hidden=%d heads=%d outside the envelope (hidden in {256,512,768,1024}, head dim 64)",
             c->hidden, c->heads);
  QF_REQUIRE(c->pooling == QF_BERT_POOL_CLS || c->pooling == QF_BERT_POOL_MEAN, QF_E_INVALID, "qf_bert: unknown pooling %d",
             c->pooling);
  return QF_OK;
}

template <int VPL>
int encode_impl(const qf_bert_config& c, const bf16* wts, const int32_t* tokens, const int32_t* lens, const int32_t* cu,
                long long M_packed, int B, int L, bf16* out_bf16, float* out_f32, int32_t* out_status, unsigned char* ws,
                cudaStream_t st) {
  const long long M = cu ? M_packed : (long long)B * L;        // packed: L = longest admissible sequence
  const Offsets o = offsets(c);
  const Ws w = ws_layout(c, M);
  bf16* X = reinterpret_cast<bf16*>(ws + w.x);
  bf16* QKV = reinterpret_cast<bf16*>(ws + w.qkv);
  bf16* CTX = reinterpret_cast<bf16*>(ws + w.ctx);
  bf16* T1 = reinterpret_cast<bf16*>(ws + w.t1);
  bf16* FF = reinterpret_cast<bf16*>(ws + w.ff);
  void* lt_ws = ws + w.lt;
  const int H = c.hidden, F = c.ffn;
  const long long Mg = gemm_rows(M);
  const unsigned tok_blocks = (unsigned)((M + 7) / 8);
  if (out_status) QF_CUDA_TRY(cudaMemsetAsync(out_status, 0, (size_t)B * sizeof(int32_t), st));
  embed_ln_kernel<VPL><<<tok_blocks, 256, 0, st>>>(tokens, wts + o.word, wts + o.pos, wts + o.type, wts + o.eg,
                                                   wts + o.eb, c.ln_eps, X, M, L, c.vocab, out_status, cu, B, c.max_pos);
  QF_CUDA_TRY(cudaGetLastError());
  const int tl = (L + 31) / 32;
  const bool use_mma = L <= 128;                    // query-length sequences: the tensor-core kernel, one warp per head
  const int lt = L <= 32 ? 32 : L <= 64 ? 64 : 128;
  const size_t attn_smem = use_mma ? (size_t)4 * 2 * lt * VSTR * sizeof(bf16)
                                   : attn_q_offset(L) + 4 * HD * sizeof(float);
  auto attn = use_mma ? (lt == 32 ? attn_mma_kernel<32> : lt == 64 ? attn_mma_kernel<64> : attn_mma_kernel<128>)
              : tl <= 8 ? attn_kernel<8> : attn_kernel<16>;
  QF_CUDA_TRY(cudaFuncSetAttribute(attn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
  const dim3 attn_grid((unsigned)B, use_mma ? (unsigned)c.heads / 4 : (unsigned)c.heads);
  // Linear layers: the own tcgen05 GEMM (csrc/qf_gemm.cu; FFN-1 with the erf-GELU fused into its epilogue) whenever the
  // shape fits its envelope (every BERT-base / -large shape does); cuBLASLt + the separate GELU pass otherwise, or with
  // QF_BERT_CUBLASLT=1 (A/B hook, read once).
  static const bool force_lt = getenv("QF_BERT_CUBLASLT") != nullptr;
  const bool own = !force_lt && qf::gemm_supported(Mg, 3 * H, H) && qf::gemm_supported(Mg, H, H) &&
                   qf::gemm_supported(Mg, F, H) && qf::gemm_supported(Mg, H, F);
  auto linear = [&](const bf16* W, const bf16* Xin, const bf16* b, bf16* Y, int N_, int K_, int act) -> int {
    if (own) return qf::gemm_bias_act(W, Xin, b, Y, Mg, N_, K_, act, st);
    return gemm_bias(W, Xin, b, Y, Mg, N_, K_, lt_ws, st);
  };
  for (int l = 0; l < c.n_layers; ++l) {
    const bf16* lw = wts + o.layer0 + (size_t)l * o.per_layer;
    int rc = linear(lw + o.wqkv, X, lw + o.bqkv, QKV, 3 * H, H, 0);
    if (rc != QF_OK) return rc;
    attn<<<attn_grid, 128, attn_smem, st>>>(QKV, lens, cu, CTX, L, H);
    QF_CUDA_TRY(cudaGetLastError());
    rc = linear(lw + o.wo, CTX, lw + o.bo, T1, H, H, 0);
    if (rc != QF_OK) return rc;
    add_ln_kernel<VPL><<<tok_blocks, 256, 0, st>>>(T1, X, lw + o.g1, lw + o.b1, c.ln_eps, X, M);
    QF_CUDA_TRY(cudaGetLastError());
    rc = linear(lw + o.w1, X, lw + o.bi1, FF, F, H, 1);            // own GEMM: GELU fused
    if (rc != QF_OK) return rc;
    if (!own) {
      const size_t n8 = (size_t)M * F / 8;
      const size_t gelu_cap = (size_t)qf::sm_count() * 32, gelu_need = (n8 + 255) / 256;
      gelu_kernel<<<(unsigned)(gelu_need < gelu_cap ? gelu_need : gelu_cap), 256, 0, st>>>(FF, n8);
      QF_CUDA_TRY(cudaGetLastError());
    }
    rc = linear(lw + o.w2, FF, lw + o.bi2, T1, H, F, 0);
    if (rc != QF_OK) return rc;
    add_ln_kernel<VPL><<<tok_blocks, 256, 0, st>>>(T1, X, lw + o.g2, lw + o.b2, c.ln_eps, X, M);
    QF_CUDA_TRY(cudaGetLastError());
  }
  pool_norm_kernel<VPL><<<(unsigned)((B + 7) / 8), 256, 0, st>>>(X, lens, cu, L, c.pooling, out_bf16, out_f32, B,
                                                                 out_status);
  QF_CUDA_TRY(cudaGetLastError());
  return QF_OK;
}

int encode_dispatch(const qf_bert_config& c, const bf16* w, const int32_t* tokens, const int32_t* lens, const int32_t* cu,
                    long long M, int B, int L, bf16* ob, float* out_f32, int32_t* out_status, unsigned char* wsb,
                    cudaStream_t st) {
  switch (c.hidden / 256) {
    case 1: return encode_impl<1>(c, w, tokens, lens, cu, M, B, L, ob, out_f32, out_status, wsb, st);
    case 2: return encode_impl<2>(c, w, tokens, lens, cu, M, B, L, ob, out_f32, out_status, wsb, st);
    case 3: return encode_impl<3>(c, w, tokens, lens, cu, M, B, L, ob, out_f32, out_status, wsb, st);
    default: return encode_impl<4>(c, w, tokens, lens, cu, M, B, L, ob, out_f32, out_status, wsb, st);
  }
}

}  // namespace

extern "C" size_t qf_bert_weight_elems(const qf_bert_config* c) {
  if (check_config(c) != QF_OK) return 0;
  return offsets(*c).total;
}

extern "C" size_t qf_bert_workspace_bytes(const qf_bert_config* c, int B, int L) {
  if (check_config(c) != QF_OK || B < 0 || L < 1) return 0;
  return ws_layout(*c, (long long)B * L).total;
}

extern "C" int qf_bert_encode(const qf_bert_config* c, const void* weights_bf16, const int32_t* tokens,
                              const int32_t* lens, int B, int L, void* out_bf16, float* out_f32, int32_t* out_status,
                              void* ws, size_t ws_bytes, void* stream) {
  int rc = check_config(c);
  if (rc != QF_OK) return rc;
  QF_REQUIRE(B >= 0 && L >= 1, QF_E_INVALID, "qf_bert_encode: bad sizes B=%d L=%d", B, L);
  QF_REQUIRE(L <= c->max_pos && L <= 512, QF_E_UNSUPPORTED, "qf_bert_encode: L=%d exceeds max_pos=%d / 512", L, c->max_pos);
  if (B == 0) return QF_OK;
  QF_REQUIRE(weights_bf16 && tokens && ws && (out_bf16 || out_f32), QF_E_INVALID, "qf_bert_encode: null pointer argument");
  QF_REQUIRE(ws_bytes >= ws_layout(*c, (long long)B * L).total, QF_E_WORKSPACE,
             "qf_bert_encode: workspace %zu < %zu bytes", ws_bytes, ws_layout(*c, (long long)B * L).total);
  rc = qf::check_device();
  if (rc != QF_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bf16* w = reinterpret_cast<const bf16*>(weights_bf16);
  bf16* ob = reinterpret_cast<bf16*>(out_bf16);
  unsigned char* wsb = reinterpret_cast<unsigned char*>(ws);
  return encode_dispatch(*c, w, tokens, lens, nullptr, 0, B, L, ob, out_f32, out_status, wsb, st);
}

extern "C" int qf_bert_encode_packed(const qf_bert_config* c, const void* weights_bf16, const int32_t* tokens,
                                     const int32_t* cu_seqlens, int B, long long M, int max_len, void* out_bf16,
                                     float* out_f32, int32_t* out_status, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_config(c);
  if (rc != QF_OK) return rc;
  QF_REQUIRE(B >= 0 && M >= 0 && max_len >= 1, QF_E_INVALID, "qf_bert_encode_packed: bad sizes B=%d M=%lld max_len=%d", B, M,
             max_len);
  QF_REQUIRE(max_len <= c->max_pos && max_len <= 512, QF_E_UNSUPPORTED,
             "qf_bert_encode_packed: max_len=%d exceeds max_pos=%d / 512", max_len, c->max_pos);
  if (B == 0) return QF_OK;
  QF_REQUIRE(M >= 1 && M <= (long long)B * max_len, QF_E_INVALID,
             "qf_bert_encode_packed: M=%lld outside [1, B * max_len = %lld]", M, (long long)B * max_len);
  QF_REQUIRE(weights_bf16 && tokens && cu_seqlens && ws && (out_bf16 || out_f32), QF_E_INVALID,
             "qf_bert_encode_packed: null pointer argument");
  QF_REQUIRE(ws_bytes >= ws_layout(*c, M).total, QF_E_WORKSPACE, "qf_bert_encode_packed: workspace %zu < %zu bytes",
             ws_bytes, ws_layout(*c, M).total);
  rc = qf::check_device();
  if (rc != QF_OK) return rc;
  return encode_dispatch(*c, reinterpret_cast<const bf16*>(weights_bf16), tokens, nullptr, cu_seqlens, M, B, max_len,
                         reinterpret_cast<bf16*>(out_bf16), out_f32, out_status, reinterpret_cast<unsigned char*>(ws),
                         reinterpret_cast<cudaStream_t>(stream));
}
